/*
 * bdrc_b200.h -- C ABI of libbdrc_b200.so: the B200 (sm_100a) page->text hot path of the
 * BDRC OCR pipeline (buda-base/ocr-inference).
 *
 * The reference has no FFI of its own for this path: the native boundary is
 * onnxruntime's InferenceSession.run_with_ort_values plus the numpy/cv2 calls around it
 * (BDRC/inference.py:128-130,152-160,294,359-370).  Every entry point below replaces one
 * such call site; the citation on each is the reference interface it stands in for.
 *
 * Conventions
 *   - plain pointers and sizes only; all data pointers are DEVICE pointers unless the name
 *     ends in _host; `stream` is a cudaStream_t passed as void*.
 *   - every function returns 0 on success or a BDRC_ERR_* code; the message is available
 *     from bdrc_last_error(ctx).  Nothing throws, nothing is allocated behind the caller's
 *     back: scratch memory is sized by the *_workspace_bytes() queries and passed in.
 *   - one bdrc_ctx per GPU; a ctx is not thread-safe (the reference objects are not
 *     either).
 *   - there is NO CPU fallback: without a CUDA device bdrc_create fails.
 */
#ifndef BDRC_B200_H
#define BDRC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BDRC_OK 0
#define BDRC_ERR_ARG 1
#define BDRC_ERR_CUDA 2
#define BDRC_ERR_UNSUPPORTED 3
#define BDRC_ERR_WEIGHTS 4

#define BDRC_MODEL_SEG_DLV3P_R34 1 /* DeepLabV3+/ResNet-34, PyTorch-LineInference.ipynb:L112-L469 */
#define BDRC_MODEL_OCR_EASTER2 2   /* Easter2 1-D conv recogniser, SURVEY 8(a)-8 */

#define BDRC_POST_SIGMOID 0 /* LineDetection.predict   BDRC/inference.py:170-183 */
#define BDRC_POST_SOFTMAX 1 /* LayoutDetection.predict BDRC/inference.py:263-276 */

#define BDRC_LAYOUT_TV 0 /* logits [L][T][V] */
#define BDRC_LAYOUT_VT 1 /* logits [L][V][T] (what the reference's ONNX emits, inference.py:373-376) */

typedef struct bdrc_ctx bdrc_ctx;
typedef struct bdrc_model bdrc_model;

/* ---- context -------------------------------------------------------------------------- */
int bdrc_create(int device, bdrc_ctx** out);
void bdrc_destroy(bdrc_ctx* ctx);
const char* bdrc_last_error(bdrc_ctx* ctx);
const char* bdrc_version(void);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
uint64_t bdrc_launch_count(bdrc_ctx* ctx);
/* Convolutions that ran on the SIMT engine instead of tcgen05.  A shape without a tcgen05 tiling is an
 * error (BDRC_ERR_UNSUPPORTED) unless BDRC_ALLOW_SIMT=1 was set when the ctx was created or the SIMT engine
 * was selected explicitly; bench.py prints this counter (0 on every measured path). */
uint64_t bdrc_simt_launch_count(bdrc_ctx* ctx);
/* tensor-map / tiling plans built so far: a steady-state forward re-uses cached plans (no host-side
 * cuTensorMapEncodeTiled per launch), so this stops growing once the buffers of a workload repeat */
uint64_t bdrc_conv_plans_built(bdrc_ctx* ctx);
/* conv engine: 0 = tcgen05/TMEM implicit GEMM (product path), 1 = SIMT (debug cross-check) */
int bdrc_set_conv_engine(bdrc_ctx* ctx, int engine);
/* decoder tail of the segmentation network: 1 = one fused kernel (default), 0 = the three-kernel form (cross-check in
 * tests; BDRC_NO_DEC_TAIL=1 selects it at context creation) */
int bdrc_set_dec_tail(bdrc_ctx* ctx, int fused);
/* 64 -> 64 channel 3x3 layers on 128-pixel-wide maps (ResNet layer 1): 1 = the wide-N kernel (three horizontal taps
 * in the N dimension, default), 0 = the slab kernel (cross-check in tests; BDRC_NO_WIDE_CONV=1 at context creation) */
int bdrc_set_wide_conv(bdrc_ctx* ctx, int on);
/* host threads waiting for this device sleep (on = 1, cudaDeviceScheduleBlockingSync) or spin (0, the driver's
 * default); applies to the device's primary context, i.e. to every runtime user in the process */
int bdrc_set_blocking_sync(bdrc_ctx* ctx, int on);
/* per-launch CUDA-event profile of the network kernels (bench.py roofline numbers) */
int bdrc_profile_enable(bdrc_ctx* ctx, int on);
int bdrc_profile_reset(bdrc_ctx* ctx);
int bdrc_profile_count(bdrc_ctx* ctx);
/* synchronises the recorded events; ms and flops are per launch, name is a static string */
int bdrc_profile_get(bdrc_ctx* ctx, int i, const char** name, float* ms, double* flops);

/* host-side staging helper (no GPU work): copies n byte ranges src[i][0..nbytes[i]) to dst + dst_off[i]
 * with up to `threads` native threads.  Used to pack ragged line crops into one pinned upload buffer:
 * the reference assembles its OCR input line by line in numpy (BDRC/inference.py:336-347, 618-666). */
int bdrc_host_gather(const void* const* src, const int64_t* nbytes, const int64_t* dst_off, int n, void* dst,
                     int threads);
/* n independent host -> host copies with up to `threads` native threads (no GPU work): hands result masks out of
 * the pinned staging buffers as fresh caller-owned arrays, the ownership the reference's predict() has
 * (BDRC/inference.py:170-183 returns a new ndarray per call). */
int bdrc_host_copy(const void* const* src, void* const* dst, const int64_t* nbytes, int n, int threads);

/* ---- models: replaces ort.InferenceSession(path)  BDRC/inference.py:128-130, 294 ------- */
/* blob_host: a "BDRCW1" weight container (see ocr_inference_b200/weights.py), fp32 OIHW
 * tensors with BN already folded (the ONNX files the reference loads are folded too).
 * n_out = number of segmentation classes, or the OCR vocabulary size V (blank included). */
int bdrc_load_weights(bdrc_ctx* ctx, int model_kind, const void* blob_host, size_t blob_bytes,
                      int n_out, bdrc_model** out);
void bdrc_free_model(bdrc_model* m);

/* ---- a-1: Detection._preprocess_image  BDRC/inference.py:132-139 ----------------------- */
/* (utils.preprocess_image :741-778, tile_image :713-721, binarize :796-815, normalize :781-793)
 * page: HWC u8 [H][W][3], already clamp-resized (bdrc_resize_bilinear_u8) if needed.
 * Pads bottom/right with 255 to a multiple of `patch`, tiles row-major, per tile
 * gray -> 51x51 Gaussian adaptive threshold (C=13, replicate border at TILE edges) and
 * writes u8 {0,1} tiles [nty*ntx][patch][patch] (the three reference channels are
 * identical, so one is stored).  nty = ceil(H/patch), ntx = ceil(W/patch). */
int bdrc_seg_preprocess(bdrc_ctx* ctx, const uint8_t* page, int H, int W, int patch,
                        uint8_t* tiles, void* stream);
/* a-1 for several pages in ONE launch (Detection._preprocess_image per page, BDRC/inference.py:132-139, as
 * OCRPipeline runs it page after page).  pages / H / W are HOST arrays of n_pages device pointers and (already
 * clamp-resized) sizes; tiles receives the tiles of all pages back to back, page after page. */
int bdrc_seg_preprocess_batch(bdrc_ctx* ctx, const uint8_t* const* pages, const int* H, const int* W, int n_pages,
                              int patch, uint8_t* tiles, void* stream);

/* torch-path collate  utils.multi_image_collate_fn :1142-1172 (no binarise): resize_clamp
 * (F.interpolate bilinear, align_corners=False, to Hc x Wc; identity when equal) -> pad with 255
 * to a multiple of `patch` -> row-major tiles fp32 [nty*ntx][patch][patch][3] with values in
 * [0,255]; the /255 happens in the network stem. */
int bdrc_seg_tile_rgb(bdrc_ctx* ctx, const uint8_t* page, int H, int W, int Hc, int Wc, int patch,
                      float* tiles, void* stream);

/* ---- a-2: Detection._predict  BDRC/inference.py:152-160 -------------------------------- */
/* tiles: cin = 1 -> u8 [N][512][512] with values {0,1} (bdrc_seg_preprocess);
 *        cin = 3 -> fp32 [N][512][512][3] with values in [0,255] (bdrc_seg_tile_rgb), scaled by
 *        1/255 in the stem.
 * logits: fp32 NCHW [N][C][512][512], like the ONNX "output". */
size_t bdrc_seg_workspace_bytes(bdrc_model* m, int n_tiles);
int bdrc_seg_forward(bdrc_ctx* ctx, bdrc_model* m, const void* tiles, int n_tiles, int cin,
                     float* logits, void* workspace, size_t workspace_bytes, void* stream);
/* same network, but stops at the 128x128 head output (fp32 NHWC [N][128][128][C]); the x4
 * align_corners upsample is then fused into bdrc_seg_postprocess_lowres. */
int bdrc_seg_forward_lowres(bdrc_ctx* ctx, bdrc_model* m, const void* tiles, int n_tiles,
                            int cin, float* logits_lowres, void* workspace,
                            size_t workspace_bytes, void* stream);

/* ---- a-3: LineDetection.predict / LayoutDetection.predict post  inference.py:175-181 ---- */
/* logits fp32 NCHW [nty*ntx][C][patch][patch] -> mask u8 [H0][W0][C] {0,255} (C=1: [H0][W0]).
 * sigmoid(x) > thr (mode SIGMOID) or softmax_c(x) > thr (mode SOFTMAX), stitch row-major,
 * crop the padding (Hc x Wc = clamped page size), and if (Hc,Wc) != (H0,W0) apply the
 * reference's cv2.resize of the {0,1} map followed by astype(u8) truncation
 * (inference.py:141-150,180): a pixel is 255 only if its bilinear value is exactly 1. */
int bdrc_seg_postprocess(bdrc_ctx* ctx, const float* logits, int nty, int ntx, int C, int patch,
                         int Hc, int Wc, int H0, int W0, int mode, float thr, uint8_t* mask,
                         void* stream);
/* fused x4 align_corners bilinear upsample (PyTorch-LineInference.ipynb:L465-L469) + the above */
int bdrc_seg_postprocess_lowres(bdrc_ctx* ctx, const float* logits_lowres, int nty, int ntx,
                                int C, int patch, int Hc, int Wc, int mode, float thr,
                                uint8_t* mask, void* stream);
/* a-3 for several pages in ONE launch (LineDetection.predict per page, BDRC/inference.py:170-183): host arrays of
 * per-page low-res logits pointers, tile counts, sizes (no page resize: Hc x Wc is the mask size) and mask pointers. */
int bdrc_seg_postprocess_lowres_batch(bdrc_ctx* ctx, const float* const* logits_lowres, const int* nty,
                                      const int* ntx, const int* Hc, const int* Wc, int n_pages, int C, int patch,
                                      int mode, float thr, uint8_t* const* masks, void* stream);

/* torch-path post  inference.py:947-956: softmax -> stitch -> crop -> float bilinear
 * (align_corners=False) to H0 x W0 -> > thr -> u8*255, planar [C][H0][W0]. */
int bdrc_seg_postprocess_torch(bdrc_ctx* ctx, const float* logits, int nty, int ntx, int C,
                               int patch, int Hc, int Wc, int H0, int W0, float thr,
                               uint8_t* mask, void* stream);

/* ---- resize / remap: cv2.resize, cv2.warpAffine, map_coordinates ----------------------- */
/* cv2.resize(INTER_LINEAR) on u8, 11-bit fixed point (utils.py:457-494, inference.py:148) */
int bdrc_resize_bilinear_u8(bdrc_ctx* ctx, const uint8_t* src, int Hs, int Ws, int C,
                            uint8_t* dst, int Hd, int Wd, void* stream);
/* F.interpolate(bilinear, align_corners=False) on u8->fp32 rounding as utils.resize_clamp
 * :967-999 does (float result, then consumed as float/255 by the stem): writes fp32 */
int bdrc_resize_bilinear_f32(bdrc_ctx* ctx, const uint8_t* src, int Hs, int Ws, int C,
                             float* dst, int Hd, int Wd, void* stream);
/* cv2.warpAffine(image, M, (W,H), INTER_LINEAR, borderValue=0)  line_detection.py:62-76.
 * M is the 2x3 forward matrix from cv2.getRotationMatrix2D (row-major, 6 doubles, host). */
int bdrc_warp_affine_u8(bdrc_ctx* ctx, const uint8_t* src, int H, int W, int C,
                        const double* M_host, uint8_t* dst, void* stream);
/* bilinear remap with an explicit coordinate field (scipy.ndimage.map_coordinates order=1,
 * mode='constant', cval=0; image_dewarping.py:66-72): coords fp32 [2][H][W] = (row, col). */
int bdrc_remap_bilinear_u8(bdrc_ctx* ctx, const uint8_t* src, int H, int W, int C,
                           const float* coords, uint8_t* dst, void* stream);
/* fused thin-plate-spline evaluate + bilinear sample (image_dewarping.py:55-72): ctrl_host
 * [n][2] control points (row,col), params_host [(n+3)][2] solved spline parameters
 * (tps.ThinPlateSpline: radial weights, then the affine rows [1, row, col]); n <= 64.
 * workspace: bdrc_tps_workspace_bytes() bytes of device memory. */
size_t bdrc_tps_workspace_bytes(void);
int bdrc_tps_remap_u8(bdrc_ctx* ctx, const uint8_t* src, int H, int W, int C,
                      const double* ctrl_host, const double* params_host, int n_ctrl,
                      uint8_t* dst, void* workspace, size_t workspace_bytes, void* stream);

/* cubic B-spline sampling, scipy.ndimage.map_coordinates(order=3, mode='constant', cval=0): what the
 * reference's run_tps asks for by calling map_coordinates with its defaults (image_dewarping.py:66-72).
 * Step 1 (scipy's spline_filter): u8 image -> fp64 spline coefficients [H][W][C], bdrc_spline3_coef_bytes()
 * bytes of device memory, each channel filtered on its own along axis 0 then axis 1 (mirror boundary).
 * Step 2: sample the coefficients at explicit coordinates (fp64 [2][H][W] = (row, col)) or at the
 * thin-plate-spline image of the output grid (same arguments as bdrc_tps_remap_u8). */
size_t bdrc_spline3_coef_bytes(int H, int W, int C);
size_t bdrc_spline3_workspace_bytes(int H, int W, int C);   /* device scratch between the two axes */
int bdrc_spline3_prefilter_u8(bdrc_ctx* ctx, const uint8_t* src, int H, int W, int C, double* coef,
                              void* workspace, size_t workspace_bytes, void* stream);
int bdrc_remap_cubic_u8(bdrc_ctx* ctx, const double* coef, int H, int W, int C, const double* coords,
                        uint8_t* dst, void* stream);
int bdrc_tps_remap_cubic_u8(bdrc_ctx* ctx, const double* coef, int H, int W, int C,
                            const double* ctrl_host, const double* params_host, int n_ctrl,
                            uint8_t* dst, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a-5: extract_line_images  BDRC/line_detection.py:330-359 --------------------------- */
/* Inputs shared by both calls: page u8 HWC [H][W][C] (C = 1 or 3, device); for each line l the
 * filled-contour mask of its bounding-box ROI roi_host[l] = (x, y, w, h) -- rasterised by the
 * caller with cv2.drawContours exactly as line_detection.py:349-351 does, but into the ROI only
 * -- packed in `masks` (device) at byte offset mask_off_host[l], row-major h x w, non-zero = set.
 *
 * bdrc_line_crop_heights: for every line and every candidate structuring element j
 * (kw_host/kh_host [L][n_cand], cv2.getStructuringElement(MORPH_RECT, (k, int(k*kf))), anchor at
 * the centre, line_detection.py:280-285) the height of the crop mask_n_crop (:79-104) would
 * produce -- what get_line_image's shrink-k loop (:311-318) looks at.  heights: device int32
 * [L][n_cand].
 *
 * bdrc_extract_lines: dilate, AND with the page, delete rows / columns in which ANY channel is
 * all zero (rows first, then columns of the remaining rows), and gather the ragged crops back
 * to back into `crops` (16-byte aligned starts): crop_hw[l] = (h, w), crop_off[l] = byte offset
 * (device arrays).  If the crops do not fit `crops_capacity` nothing is gathered (crop_hw is
 * still valid, so the caller can size the buffer and call again).
 *
 * workspace: bdrc_extract_lines_workspace_bytes(H, W, n_items, sat_elems) with n_items = L *
 * n_cand (or L) and sat_elems = sum_l (w_l + 1) * (h_l + 1). */
size_t bdrc_extract_lines_workspace_bytes(int H, int W, int n_items, int64_t sat_elems);
int bdrc_line_crop_heights(bdrc_ctx* ctx, const uint8_t* page, int H, int W, int C,
                           const uint8_t* masks, const int32_t* roi_host,
                           const int64_t* mask_off_host, int L, int n_cand,
                           const int32_t* kw_host, const int32_t* kh_host, int32_t* heights,
                           void* workspace, size_t workspace_bytes, void* stream);
int bdrc_extract_lines(bdrc_ctx* ctx, const uint8_t* page, int H, int W, int C,
                       const uint8_t* masks, const int32_t* roi_host,
                       const int64_t* mask_off_host, int L, const int32_t* kw_host,
                       const int32_t* kh_host, uint8_t* crops, size_t crops_capacity,
                       int32_t* crop_hw, int64_t* crop_off, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ---- a-10 (GPU half): connected-component census of the line mask ----------------------- */
/* The pixel-scanning part of the host geometry: cv2.boundingRect(mask) and the slice census of
 * get_line_threshold (BDRC/line_detection.py:362-421: ~70 cv2.findContours calls on 20-pixel slices whose only
 * output is "median contour centre // contour count" of the fullest slice), by union-find connected-component
 * labelling on the device: foreground components (8-connected) + holes (4-connected background components that
 * do not reach the slice frame) = the contours findContours(RETR_LIST) returns; their bounding boxes give the
 * centres.  Polygons (approxPolyDP, hulls, filled ROI masks) stay cv2 work on the host.
 * row_any: u8 [H] (1 where the row holds foreground: the host's row compaction starts from it);
 * header: int32 [16] = {x, y, w, h of boundingRect (w = 0: empty), x_steps, strip w, strip h, best slice,
 * its contour count n, centres written, overflow flag, 0...}; centers: int32 [bdrc_census_max_centers()], the
 * "cy + ch // 2" values of the best slice in unspecified order (their median is what is used).
 * All outputs are device pointers. */
size_t bdrc_mask_census_workspace_bytes(int H, int W, int slice_width);
int bdrc_census_max_centers(void);
int bdrc_mask_census(bdrc_ctx* ctx, const uint8_t* mask, int H, int W, int slice_width, uint8_t* row_any,
                     int32_t* header, int32_t* centers, void* workspace, size_t workspace_bytes, void* stream);
/* 8-connected foreground components of a mask with bounding box and pixel count
 * (cv2.connectedComponentsWithStats, connectivity 8, without the background label): n_comp (device int32) and up to
 * max_comp records int32 {x0, y0, x1, y1 (inclusive), area, root pixel} in unspecified order. */
size_t bdrc_mask_components_workspace_bytes(int H, int W);
int bdrc_mask_components(bdrc_ctx* ctx, const uint8_t* mask, int H, int W, int32_t* n_comp, int32_t* stats,
                         int max_comp, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a-7: OCRInference._prepare_ocr_line  BDRC/inference.py:313-357 --------------------- */
/* crops: L ragged HWC u8 (3 ch) images at byte offsets crop_off, sizes crop_hw (h,w).
 * Per line: optional h x h white pre-pad left/right, aspect-preserving cv2 resize to fit
 * Wt x Ht, centre pad with black, adaptive-threshold binarise, -> {-1,+1}.
 * Output `lines`: bf16 [L][Wt][Cpad] time-major (image rows are channels, zero-padded from
 * Ht to Cpad; EVERY element is written, the buffer needs no initialisation) -- the layout the
 * recogniser's first conv consumes.  `lines_f32` (optional, may be NULL) receives the
 * reference layout fp32 [L][Ht][Wt] for parity checks.  `lines` may be NULL when `lines_f32`
 * is given (then any Ht is accepted; with `lines`, Cpad >= Ht). */
size_t bdrc_ocr_prepare_workspace_bytes(int L, int Ht, int Wt);
int bdrc_ocr_prepare_lines(bdrc_ctx* ctx, const uint8_t* crops, const int32_t* crop_hw,
                           const int64_t* crop_off, int L, int Ht, int Wt, int pre_pad,
                           int Cpad, void* lines_bf16, float* lines_f32, void* workspace,
                           size_t workspace_bytes, void* stream);

/* ---- a-8: OCRInference._predict  BDRC/inference.py:359-370 ------------------------------ */
/* lines: bf16 [L][Wt][Cpad] from bdrc_ocr_prepare_lines; logits fp32 [L][T][V] (T = Wt/4). */
size_t bdrc_ocr_workspace_bytes(bdrc_model* m, int L, int Wt);
int bdrc_ocr_forward(bdrc_ctx* ctx, bdrc_model* m, const void* lines_bf16, int L, int Wt,
                     float* logits, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a-9: softmax + CTC greedy decode  BDRC/inference.py:372-384, 108-112 --------------- */
/* per timestep argmax (ties -> lowest index, as np.argmax), collapse repeats, drop `blank`.
 * ids int32 [L][T] (first lens[l] entries valid), lens int32 [L]. */
int bdrc_ctc_greedy(bdrc_ctx* ctx, const float* logits, int L, int V, int T, int layout,
                    int blank, int32_t* ids, int32_t* lens, void* stream);

/* Front end of the beam-search decode (CTCDecoder.ctc_beam_decode / pyctcdecode decode_beams, BDRC/inference.py:
 * 114-118, 386-414): per (line, time step) the candidate tokens of the log-softmaxed row -- log p >= min_logp
 * (pyctcdecode's token_min_logp, -5), at most bdrc_ctc_topk_width() of them, sorted by descending log p (ties: lower
 * index), the arg-max always present.  tokens / logps: [L][T][width], counts: [L][T].  The prefix merge over
 * these candidates is host code (ocr_inference_b200/ctc_beam.py). */
int bdrc_ctc_topk(bdrc_ctx* ctx, const float* logits, int L, int V, int T, int layout, float min_logp,
                  int32_t* tokens, float* logps, int32_t* counts, void* stream);
int bdrc_ctc_topk_width(void);

/* ---- test hook: one convolution through the implicit-GEMM engine ------------------------ */
/* x bf16 NHWC [N][H][W][Cin]; w fp32 OIHW (host); bias fp32 (host, may be NULL); residual
 * bf16 NHWC or NULL; y bf16 (or fp32 if out_fp32) NHWC [N][Ho][Wo][Cout]. */
int bdrc_conv2d_test(bdrc_ctx* ctx, const void* x, int N, int H, int W, int Cin,
                     const float* w_host, const float* bias_host, int Cout, int R, int S,
                     int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                     const void* residual, int relu, int out_fp32, void* y, void* stream);
/* Test hook: depthwise 3x3 (pad = dilation, no bias) on bf16 NHWC x [N][H][W][C], weights fp32 [9][C] on the
 * host, y like x; synchronises the stream.  Used by tools/ and tests to exercise the kernel variants directly. */
int bdrc_depthwise3x3_test(bdrc_ctx* ctx, const void* x, int N, int H, int W, int C, int dil,
                           const float* w9c_host, void* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BDRC_B200_H */
