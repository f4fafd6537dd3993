// HBM-bound pre/post-processing stages of the page->text path (everything around the two
// networks): tiling + adaptive-threshold binarisation, mask thresholding / stitching / resize,
// cv2-exact bilinear resize, OCR line preparation and greedy CTC decode.
#include "adaptive.cuh"
#include "tc_ptx.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// cv2.resize(INTER_LINEAR) coordinate tables, computed on the fly exactly as OpenCV does:
//   scale = 1.0 / ((double)dsize / ssize);  fx = (float)((d + 0.5) * scale - 0.5);
//   sx = floor(fx); fx -= sx; horizontal: fold out-of-range taps; vertical: clip index only.
// (oracle/hostops.py:_linear_coefs)
// ------------------------------------------------------------------------------------------
struct LinCoef {
  int i0, i1;
  float w0, w1;
};

// `scale` = 1.0 / ((double)dsize / (double)ssize), hoisted by callers that evaluate many pixels of one resize
__device__ __forceinline__ LinCoef lin_coef_scaled(int d, int ssize, double scale, bool vertical) {
  float fx = (float)__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
  int sx = (int)floorf(fx);
  fx = fx - (float)sx;
  if (!vertical) {
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= ssize - 1) { fx = 0.f; sx = ssize - 1; }
  }
  LinCoef c;
  c.i0 = min(max(sx, 0), ssize - 1);
  c.i1 = min(max(sx + 1, 0), ssize - 1);
  c.w0 = 1.0f - fx;
  c.w1 = fx;
  return c;
}

__device__ __forceinline__ LinCoef lin_coef(int d, int ssize, int dsize, bool vertical) {
  const double scale = 1.0 / ((double)dsize / (double)ssize);
  float fx = (float)__dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
  int sx = (int)floorf(fx);
  fx = fx - (float)sx;
  if (!vertical) {
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= ssize - 1) { fx = 0.f; sx = ssize - 1; }
  }
  LinCoef c;
  c.i0 = min(max(sx, 0), ssize - 1);
  c.i1 = min(max(sx + 1, 0), ssize - 1);
  c.w0 = 1.0f - fx;
  c.w1 = fx;
  return c;
}

// CV_64F variant: the fraction stays in double (see oracle/hostops.py:_linear_coefs_f64)
struct LinCoef64 {
  int i0, i1;
  double f;
};

__device__ __forceinline__ LinCoef64 lin_coef64(int d, int ssize, int dsize, bool vertical) {
  const double scale = 1.0 / ((double)dsize / (double)ssize);
  // explicit rn intrinsics: no FMA contraction, same roundings as the numpy oracle / OpenCV
  double fx = __dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
  int sx = (int)floor(fx);
  fx = fx - (double)sx;
  if (!vertical) {
    if (sx < 0) { fx = 0.0; sx = 0; }
    if (sx >= ssize - 1) { fx = 0.0; sx = ssize - 1; }
  }
  LinCoef64 c;
  c.i0 = min(max(sx, 0), ssize - 1);
  c.i1 = min(max(sx + 1, 0), ssize - 1);
  c.f = fx;
  return c;
}

__device__ __forceinline__ int fix11(float w) { return __float2int_rn(w * 2048.0f); }

// one output sample of the 11-bit fixed-point bilinear (VResizeLinear 8u SIMD formula)
__device__ __forceinline__ int resize_fix_sample(int s00, int s01, int s10, int s11, int a0, int a1,
                                                 int b0, int b1) {
  const int r0 = s00 * a0 + s01 * a1;
  const int r1 = s10 * a0 + s11 * a1;
  int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
  return min(max(v, 0), 255);
}

__global__ void resize_u8_kernel(const uint8_t* __restrict__ src, int Hs, int Ws, int C,
                                 uint8_t* __restrict__ dst, int Hd, int Wd) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= Wd) return;
  uint8_t* o = dst + ((size_t)y * Wd + x) * C;
  if (Hs == Hd && Ws == Wd) {
    for (int c = 0; c < C; ++c) o[c] = src[((size_t)y * Ws + x) * C + c];
    return;
  }
  if (Ws == 2 * Wd && Hs == 2 * Hd) {   // cv2's INTER_AREA fast path for exact 1/2 scaling
    const uint8_t* p = src + ((size_t)(2 * y) * Ws + 2 * x) * C;
    for (int c = 0; c < C; ++c)
      o[c] = (uint8_t)((p[c] + p[C + c] + p[(size_t)Ws * C + c] + p[(size_t)Ws * C + C + c] + 2) >> 2);
    return;
  }
  const LinCoef cx = lin_coef(x, Ws, Wd, false), cy = lin_coef(y, Hs, Hd, true);
  const int a0 = fix11(cx.w0), a1 = fix11(cx.w1), b0 = fix11(cy.w0), b1 = fix11(cy.w1);
  const uint8_t* r0 = src + (size_t)cy.i0 * Ws * C;
  const uint8_t* r1 = src + (size_t)cy.i1 * Ws * C;
  for (int c = 0; c < C; ++c)
    o[c] = (uint8_t)resize_fix_sample(r0[cx.i0 * C + c], r0[cx.i1 * C + c], r1[cx.i0 * C + c],
                                      r1[cx.i1 * C + c], a0, a1, b0, b1);
}

// torch F.interpolate(bilinear, align_corners=False) u8 -> fp32 (utils.resize_clamp :990-996)
__global__ void resize_f32_kernel(const uint8_t* __restrict__ src, int Hs, int Ws, int C,
                                  float* __restrict__ dst, int Hd, int Wd) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= Wd) return;
  const float sy = (float)Hs / (float)Hd, sx = (float)Ws / (float)Wd;
  float fy = sy * ((float)y + 0.5f) - 0.5f;
  float fx = sx * ((float)x + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy;
  fx = fx < 0.f ? 0.f : fx;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < Hs - 1 ? 1 : 0), x1 = x0 + (x0 < Ws - 1 ? 1 : 0);
  const float ly1 = fy - (float)y0, lx1 = fx - (float)x0;
  const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  for (int c = 0; c < C; ++c) {
    const float v00 = src[((size_t)y0 * Ws + x0) * C + c], v01 = src[((size_t)y0 * Ws + x1) * C + c];
    const float v10 = src[((size_t)y1 * Ws + x0) * C + c], v11 = src[((size_t)y1 * Ws + x1) * C + c];
    dst[((size_t)y * Wd + x) * C + c] = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
  }
}

// ------------------------------------------------------------------------------------------
// torch-path collate (utils.resize_clamp / pad_to_multiple / tile_timage :967-1023): float
// bilinear (align_corners=False) resize, pad 255, RGB fp32 tiles with values in [0,255]
// ------------------------------------------------------------------------------------------
__global__ void tile_rgb_kernel(const uint8_t* __restrict__ page, int H, int W, int Hc, int Wc,
                                int patch, int ntx, float* __restrict__ tiles) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;   // pixel column inside the tile
  const int y = blockIdx.y;                                // tile row
  const int t = blockIdx.z;                                // tile
  if (x >= patch) return;
  const int ty = t / ntx, tx = t - ty * ntx;
  const int py = ty * patch + y, px = tx * patch + x;
  float* o = tiles + (((size_t)t * patch + y) * patch + x) * 3;
  if (py >= Hc || px >= Wc) { o[0] = o[1] = o[2] = 255.f; return; }
  if (Hc == H && Wc == W) {
    const uint8_t* p = page + ((size_t)py * W + px) * 3;
    o[0] = p[0]; o[1] = p[1]; o[2] = p[2];
    return;
  }
  const float sy = (float)H / (float)Hc, sx = (float)W / (float)Wc;
  float fy = sy * ((float)py + 0.5f) - 0.5f, fx = sx * ((float)px + 0.5f) - 0.5f;
  fy = fy < 0.f ? 0.f : fy; fx = fx < 0.f ? 0.f : fx;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
  const float ly1 = fy - (float)y0, lx1 = fx - (float)x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  for (int c = 0; c < 3; ++c) {
    const float v00 = page[((size_t)y0 * W + x0) * 3 + c], v01 = page[((size_t)y0 * W + x1) * 3 + c];
    const float v10 = page[((size_t)y1 * W + x0) * 3 + c], v11 = page[((size_t)y1 * W + x1) * 3 + c];
    o[c] = ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11);
  }
}

// ------------------------------------------------------------------------------------------
// a-3 mask thresholding
// ------------------------------------------------------------------------------------------
struct PostParams {
  const float* logits;   // [nt][C][P][P]  (or low-res [nt][C][P/4][P/4])
  int ntx, C, P, lowres;
  int Hc, Wc, H0, W0;
  int mode;
  float thr, thr_logit;
  uint8_t* mask;         // [H0][W0][C]
};

// logit at padded-page pixel (py, px) for class c
__device__ __forceinline__ float logit_at(const PostParams& p, int py, int px, int c) {
  const int ty = py / p.P, tx = px / p.P;
  const int y = py - ty * p.P, x = px - tx * p.P;
  const int t = ty * p.ntx + tx;
  if (!p.lowres) return p.logits[(((size_t)t * p.C + c) * p.P + y) * p.P + x];
  // x4 bilinear, align_corners=True, fp32 like torch upsample_bilinear2d
  const int L = p.P / 4;
  const float sc = (float)(L - 1) / (float)(p.P - 1);
  const float fy = sc * (float)y, fx = sc * (float)x;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < L - 1 ? 1 : 0), x1 = x0 + (x0 < L - 1 ? 1 : 0);
  const float ly1 = fy - (float)y0, lx1 = fx - (float)x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const float* b = p.logits + (size_t)t * L * L * p.C + c;   // low-res logits are NHWC
  return bdrc_bilerp(b[(y0 * L + x0) * p.C], b[(y0 * L + x1) * p.C], b[(y1 * L + x0) * p.C], b[(y1 * L + x1) * p.C],
                     lx0, lx1, ly0, ly1);
}

// {0,1} decisions for all classes of one padded-page pixel (bit c of the result)
__device__ __forceinline__ uint32_t decide(const PostParams& p, int py, int px) {
  if (p.mode == BDRC_POST_SIGMOID) {
    uint32_t m = 0;
    for (int c = 0; c < p.C; ++c) m |= (logit_at(p, py, px, c) >= p.thr_logit ? 1u : 0u) << c;
    return m;
  }
  // softmax over classes in fp32, as scipy.special.softmax on the fp32 array (inference.py:269)
  float v[32];
  float mx = -INFINITY;
  for (int c = 0; c < p.C; ++c) { v[c] = logit_at(p, py, px, c); mx = fmaxf(mx, v[c]); }
  float s = 0.f;
  for (int c = 0; c < p.C; ++c) { v[c] = expf(v[c] - mx); s += v[c]; }
  uint32_t m = 0;
  for (int c = 0; c < p.C; ++c) m |= ((v[c] / s) > p.thr ? 1u : 0u) << c;
  return m;
}

__global__ void seg_post_kernel(const PostParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= p.W0) return;
  uint32_t m;
  if (p.Hc == p.H0 && p.Wc == p.W0) {
    m = decide(p, y, x);
  } else {
    // cv2.resize of the float64 {0,1} map then astype(u8) (inference.py:148,180): the bilinear
    // value survives the truncating cast only where it is exactly 1, i.e. where every tap with a
    // non-zero weight is 1.  cv2 keeps the fraction in double for CV_64F data
    // (oracle/hostops.py:resize_linear_f64, pinned against cv2).
    if (p.C == 1 || p.C == 3 || p.C == 4) {
      // 1/3/4-channel CV_64F images take cv2's IPP route: double fraction, lerp form
      const LinCoef64 cx = lin_coef64(x, p.Wc, p.W0, false), cy = lin_coef64(y, p.Hc, p.H0, true);
      const uint32_t all = (1u << p.C) - 1u;
      const uint32_t m00 = decide(p, cy.i0, cx.i0);
      const uint32_t m01 = cx.f == 0.0 ? all : decide(p, cy.i0, cx.i1);
      m = m00 & m01;
      if (cy.f != 0.0) {
        const uint32_t m10 = decide(p, cy.i1, cx.i0);
        const uint32_t m11 = cx.f == 0.0 ? all : decide(p, cy.i1, cx.i1);
        m &= m10 & m11;
      }
    } else {
      // other channel counts (the 5-class layout map) take OpenCV's generic route: float32
      // weights, double accumulation, value >= 1.0 survives the cast
      const LinCoef cx = lin_coef(x, p.Wc, p.W0, false), cy = lin_coef(y, p.Hc, p.H0, true);
      const uint32_t m00 = decide(p, cy.i0, cx.i0), m01 = decide(p, cy.i0, cx.i1);
      const uint32_t m10 = decide(p, cy.i1, cx.i0), m11 = decide(p, cy.i1, cx.i1);
      m = 0;
      for (int c = 0; c < p.C; ++c) {
        const double r0 = __dadd_rn(__dmul_rn((double)((m00 >> c) & 1), (double)cx.w0),
                                    __dmul_rn((double)((m01 >> c) & 1), (double)cx.w1));
        const double r1 = __dadd_rn(__dmul_rn((double)((m10 >> c) & 1), (double)cx.w0),
                                    __dmul_rn((double)((m11 >> c) & 1), (double)cx.w1));
        const double v = __dadd_rn(__dmul_rn(r0, (double)cy.w0), __dmul_rn(r1, (double)cy.w1));
        m |= (v >= 1.0 ? 1u : 0u) << c;
      }
    }
  }
  uint8_t* o = p.mask + ((size_t)y * p.W0 + x) * p.C;
  for (int c = 0; c < p.C; ++c) o[c] = ((m >> c) & 1) ? 255 : 0;
}

// fast path of the common case (line model, low-res fp32 logits [nt][P/4][P/4], no page resize): 16 consecutive
// mask bytes per thread and one 16-byte store (the generic kernel writes single bytes).  Same arithmetic, in the
// same expression, as logit_at(): the decisions are identical.
__global__ void seg_post_line_lowres_kernel(const PostParams p) {
  const int x16 = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
  const int y = blockIdx.y;
  if (x16 >= p.W0) return;
  uint32_t bits = 0;
  const int n = min(16, p.W0 - x16);
  for (int j = 0; j < n; ++j) bits |= (logit_at(p, y, x16 + j, 0) >= p.thr_logit ? 1u : 0u) << j;
  uint8_t* o = p.mask + (size_t)y * p.W0 + x16;
  if (n == 16 && (p.W0 & 15) == 0) {
    uint4 v;
    uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t b = (bits >> (4 * q)) & 15u;
      w[q] = ((b & 1u) ? 0x000000FFu : 0u) | ((b & 2u) ? 0x0000FF00u : 0u) | ((b & 4u) ? 0x00FF0000u : 0u) |
             ((b & 8u) ? 0xFF000000u : 0u);
    }
    *reinterpret_cast<uint4*>(o) = v;
  } else {
    for (int j = 0; j < n; ++j) o[j] = ((bits >> j) & 1u) ? 255 : 0;
  }
}

// The line model's mask for SEVERAL pages in one launch (bdrc_seg_postprocess_lowres_batch): low-res fp32 logits
// [tiles][P/4][P/4] of each page -> u8 {0,255} [H0][W0], x4 align_corners upsample fused.  A thread owns 16
// consecutive pixels of a row (one tile: 16 | P).  It first looks at the <= 6 x 2 low-res corners those pixels
// interpolate between: an interpolated value is a convex combination of its corners (up to rounding), so when
// every corner is above the threshold by more than a rounding margin all 16 decisions are 1, when every
// corner is below by the margin they are 0 -- on a text page that settles > 95 % of the threads with 12 loads
// and no per-pixel arithmetic.  Only threads whose corners straddle the threshold evaluate the pixels, with
// bdrc_bilerp like logit_at(): the decisions are those of the per-page kernels.
struct PostPage {
  const float* logits;
  uint8_t* mask;
  int ntx, H0, W0;
};
constexpr int POST_MAX_PAGES = 32;
struct PostBatch {
  PostPage pg[POST_MAX_PAGES];
  int P;
  float thr_logit;
};

__global__ void __launch_bounds__(256) seg_post_line_lowres_batch_kernel(const PostBatch b) {
  // block = 64 units x 4 row quads, grid = (units / 64, rows / 16, pages): no index division anywhere.
  // A thread owns 16 pixels of FOUR consecutive rows (4 | P: one tile): the four rows interpolate between at most
  // three low-res rows, so 3 x 6 corner loads settle 64 pixels.  ncu (profiles/r02b): the kernel is bound by the
  // instructions of the per-pixel path, not by bytes -- on logits with sigma = 0.5 noise a fifth of the rows straddle
  // the threshold and evaluate 16 x ~30 instructions; rows away from band edges cost ~40.
  const PostPage pg = b.pg[blockIdx.z];
  const int yq = (blockIdx.y * 4 + threadIdx.y) * 4;
  const int x16 = (blockIdx.x * 64 + threadIdx.x) * 16;
  if (yq >= pg.H0 || x16 >= pg.W0) return;
  const int P = b.P, L = P / 4;
  const int ty = yq / P, yy0 = yq - ty * P, tx = x16 / P, xx0 = x16 - tx * P;
  const float sc = (float)(L - 1) / (float)(P - 1);
  const float* t = pg.logits + (size_t)(ty * pg.ntx + tx) * L * L;
  const int n = min(16, pg.W0 - x16);
  const int xa = (int)(sc * (float)xx0);
  const int xb = min((int)(sc * (float)(xx0 + n - 1)) + 1, L - 1);
  const int ya = (int)(sc * (float)yy0);
  // per low-res row ya + r: min / max of the corners xa .. xb, and whether they are all ordered (no NaN)
  float mn[3], mx[3];
  bool ord[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float* rp = t + min(ya + r, L - 1) * L;
    // 16 pixels span at most 6 low-res columns (16 * 127 / 511 < 4): six unconditional loads with the index clamped
    // to xb, so that all 18 loads of the thread are in flight together (a loop with run-time bounds issued them one
    // L2 round trip after the other)
    float v[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = __ldg(rp + min(xa + k, xb));
    float lo = v[0], hi = v[0];
    bool o = v[0] == v[0];
#pragma unroll
    for (int k = 1; k < 6; ++k) {
      o = o && (v[k] == v[k]);
      lo = fminf(lo, v[k]);
      hi = fmaxf(hi, v[k]);
    }
    mn[r] = lo; mx[r] = hi; ord[r] = o;
  }
  const float thr = b.thr_logit;
  const bool wide = n == 16 && (pg.W0 & 15) == 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int y = yq + i;
    if (y >= pg.H0) break;
    const float fy = sc * (float)(yy0 + i);
    const int y0 = (int)fy, y1 = y0 + (y0 < L - 1 ? 1 : 0);
    const int ra = y0 - ya, rb = y1 - ya;                 // 0 / 1 and ra / ra + 1
    const float lo = fminf(ra == 0 ? mn[0] : mn[1], rb == 0 ? mn[0] : rb == 1 ? mn[1] : mn[2]);
    const float hi = fmaxf(ra == 0 ? mx[0] : mx[1], rb == 0 ? mx[0] : rb == 1 ? mx[1] : mx[2]);
    const bool ordered = (ra == 0 ? ord[0] : ord[1]) && (rb == 0 ? ord[0] : rb == 1 ? ord[1] : ord[2]);
    const float margin = 1e-5f * fmaxf(fmaxf(fabsf(lo), fabsf(hi)), fabsf(thr)) + 1e-30f;
    uint32_t bits;
    if (ordered && lo >= thr + margin && hi < INFINITY) {
      bits = 0xFFFFu;
    } else if (ordered && hi < thr - margin && lo > -INFINITY) {
      bits = 0u;
    } else {
      const float ly1 = fy - (float)y0, ly0 = 1.f - ly1;
      const float* r0 = t + y0 * L;
      const float* r1 = t + y1 * L;
      bits = 0u;
      for (int j = 0; j < n; ++j) {
        const float fx = sc * (float)(xx0 + j);
        const int x0 = (int)fx, x1 = x0 + (x0 < L - 1 ? 1 : 0);
        const float lx1 = fx - (float)x0, lx0 = 1.f - lx1;
        const float v = bdrc_bilerp(__ldg(r0 + x0), __ldg(r0 + x1), __ldg(r1 + x0), __ldg(r1 + x1), lx0, lx1, ly0, ly1);
        bits |= (v >= thr ? 1u : 0u) << j;
      }
    }
    uint8_t* o = pg.mask + (size_t)y * pg.W0 + x16;
    if (wide) {
      uint4 v;
      uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t c = (bits >> (4 * q)) & 15u;
        w[q] = ((c & 1u) ? 0x000000FFu : 0u) | ((c & 2u) ? 0x0000FF00u : 0u) | ((c & 4u) ? 0x00FF0000u : 0u) |
               ((c & 8u) ? 0xFF000000u : 0u);
      }
      *reinterpret_cast<uint4*>(o) = v;
    } else {
      for (int j = 0; j < n; ++j) o[j] = ((bits >> j) & 1u) ? 255 : 0;
    }
  }
}

// fast path: sigmoid, one class, full-res logits, no resize: 4 pixels per thread
__global__ void seg_post_line_fast_kernel(const float* __restrict__ logits, int ntx, int P, int H0,
                                          int W0, float thr_logit, uint8_t* __restrict__ mask) {
  const int x4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int y = blockIdx.y;
  if (x4 >= W0) return;
  const int ty = y / P, tx = x4 / P;
  const float* src = logits + (((size_t)(ty * ntx + tx)) * P + (y - ty * P)) * P + (x4 - tx * P);
  const float4 v = __ldg(reinterpret_cast<const float4*>(src));
  uint8_t r[4] = {(uint8_t)(v.x >= thr_logit ? 255 : 0), (uint8_t)(v.y >= thr_logit ? 255 : 0),
                  (uint8_t)(v.z >= thr_logit ? 255 : 0), (uint8_t)(v.w >= thr_logit ? 255 : 0)};
  uint8_t* o = mask + (size_t)y * W0 + x4;
  if (x4 + 3 < W0 && (W0 & 3) == 0) {
    *reinterpret_cast<uint32_t*>(o) = *reinterpret_cast<const uint32_t*>(r);
  } else {
    for (int j = 0; j < 4 && x4 + j < W0; ++j) o[j] = r[j];
  }
}

// torch-path post (inference.py:947-956): softmax -> (stitch, crop) -> float bilinear
// align_corners=False to H0 x W0 -> > thr -> planar u8 [C][H0][W0]
__global__ void seg_post_torch_kernel(const PostParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= p.W0) return;
  float prob[4][32];
  int ys[2], xs[2];
  float ly1 = 0.f, lx1 = 0.f;
  if (p.Hc == p.H0 && p.Wc == p.W0) {
    ys[0] = ys[1] = y; xs[0] = xs[1] = x;
  } else {
    const float sy = (float)p.Hc / (float)p.H0, sx = (float)p.Wc / (float)p.W0;
    float fy = sy * ((float)y + 0.5f) - 0.5f, fx = sx * ((float)x + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy; fx = fx < 0.f ? 0.f : fx;
    ys[0] = (int)fy; xs[0] = (int)fx;
    ys[1] = ys[0] + (ys[0] < p.Hc - 1 ? 1 : 0); xs[1] = xs[0] + (xs[0] < p.Wc - 1 ? 1 : 0);
    ly1 = fy - (float)ys[0]; lx1 = fx - (float)xs[0];
  }
  for (int a = 0; a < 2; ++a)
    for (int b = 0; b < 2; ++b) {
      float* v = prob[a * 2 + b];
      float mx = -INFINITY;
      for (int c = 0; c < p.C; ++c) { v[c] = logit_at(p, ys[a], xs[b], c); mx = fmaxf(mx, v[c]); }
      float s = 0.f;
      for (int c = 0; c < p.C; ++c) { v[c] = expf(v[c] - mx); s += v[c]; }
      for (int c = 0; c < p.C; ++c) v[c] = v[c] / s;
    }
  const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  for (int c = 0; c < p.C; ++c) {
    const float v = ly0 * (lx0 * prob[0][c] + lx1 * prob[1][c]) + ly1 * (lx0 * prob[2][c] + lx1 * prob[3][c]);
    p.mask[((size_t)c * p.H0 + y) * p.W0 + x] = v > p.thr ? 255 : 0;
  }
}

// smallest fp32 logit whose fp32 sigmoid exceeds fp32(thr): the reference predicate
// sigmoid(x) > thr (utils.py:563-573, inference.py:177) moved to the logit domain.
float sigmoid_logit_threshold(float thr) {
  auto pred = [thr](float x) { return 1.0f / (1.0f + expf(-x)) > thr; };
  float lo = -30.f, hi = 30.f;
  if (pred(lo)) return -INFINITY;
  if (!pred(hi)) return INFINITY;
  for (;;) {
    const float mid = (float)(((double)lo + (double)hi) / 2.0);
    if (mid == lo || mid == hi) break;
    if (pred(mid)) hi = mid; else lo = mid;
  }
  return hi;
}

// ------------------------------------------------------------------------------------------
// a-7 OCR line preparation, stage 1: (pre-pad) + aspect-preserving cv2 resize + black padding
// + gray, written as a u8 gray canvas [L][Ht][Wt]
// ------------------------------------------------------------------------------------------
constexpr int CANVAS_ROWS = 16;  // canvas rows per block
__global__ void __launch_bounds__(128)
ocr_canvas_kernel(const uint8_t* __restrict__ crops, const int32_t* __restrict__ hw,
                  const int64_t* __restrict__ off, int Ht, int Wt, int pre_pad,
                  uint8_t* __restrict__ canvas, int* __restrict__ status) {
  // Per-line constants -- the double-precision geometry of _pad_ocr_line -- and the vertical resize coefficients
  // of this block's CANVAS_ROWS rows are computed once per block, one row per thread (fp64 divisions per pixel
  // dominated the first version, a serial per-block preamble for ONE row the second: 82 k tiny blocks); the
  // horizontal coefficients of a column are computed once per thread and reused for all rows of the block.
  __shared__ int g_i[4];                       // rh, rw, top, left
  __shared__ int r_i0[CANVAS_ROWS], r_i1[CANVAS_ROWS];
  __shared__ int r_b0[CANVAS_ROWS], r_b1[CANVAS_ROWS];      // 11-bit fixed-point vertical weights
  __shared__ double g_scale_x;
  tcptx::griddep_launch();   // programmatic dependent launch inside the OCR line chain (bdrc_ocr_prepare_lines)
  tcptx::griddep_wait();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y0 = blockIdx.y * CANVAS_ROWS;
  const int l = blockIdx.z;
  const int h = hw[2 * l], w0 = hw[2 * l + 1];
  const int w = pre_pad ? w0 + 2 * h : w0;     // _pre_pad hstacks an h x h white patch each side
  if (threadIdx.x < CANVAS_ROWS) {
    const int y = y0 + threadIdx.x;
    int rh = 0, rw = 0, top = 0, left = 0;
    LinCoef cy = {0, 0, 1.f, 0.f};
    double scale_x = 1.0;
    if (h > 0 && w0 > 0) {
      // geometry of _pad_ocr_line / pad_to_width / pad_to_height in Python float (double) math
      const double wr = (double)Wt / (double)w, hr = (double)Ht / (double)h;
      if (wr <= hr) { rw = Wt; rh = (int)((double)h * wr); top = (Ht - rh) / 2; left = 0; }
      else { rh = Ht; rw = (int)((double)w * hr); top = 0; left = (Wt - rw) / 2; }
      if (rh > 0 && rw > 0) {
        scale_x = 1.0 / ((double)rw / (double)w);
        if (y >= top && y < top + rh) cy = lin_coef(y - top, h, rh, true);
      }
    }
    r_i0[threadIdx.x] = cy.i0; r_i1[threadIdx.x] = cy.i1;
    r_b0[threadIdx.x] = fix11(cy.w0); r_b1[threadIdx.x] = fix11(cy.w1);
    if (threadIdx.x == 0) {
      g_i[0] = rh; g_i[1] = rw; g_i[2] = top; g_i[3] = left;
      g_scale_x = scale_x;
    }
  }
  __syncthreads();
  if (x >= Wt) return;
  const int rh = g_i[0], rw = g_i[1], top = g_i[2], left = g_i[3];
  const bool live = h > 0 && w0 > 0;
  if (x == 0 && y0 == 0) status[l] = (live && (rh <= 0 || rw <= 0)) ? 1 : 0;   // 1: cv2.resize would raise on an empty target
  const bool in_x = live && rh > 0 && rw > 0 && x >= left && x < left + rw;
  const int dx = x - left;
  const uint8_t* base = crops + off[l];
  auto fetch = [&](int sy, int sx, int* out3) {
    if (pre_pad && (sx < h || sx >= h + w0)) { out3[0] = out3[1] = out3[2] = 255; return; }
    const uint8_t* p = base + ((size_t)sy * w0 + (pre_pad ? sx - h : sx)) * 3;
    out3[0] = p[0]; out3[1] = p[1]; out3[2] = p[2];
  };
  const int mode = !in_x ? 0 : (rh == h && rw == w) ? 1 : (w == 2 * rw && h == 2 * rh) ? 2 : 3;
  LinCoef cx = {0, 0, 1.f, 0.f};
  int a0 = 0, a1 = 0;
  int o0 = 0, o1 = 0;            // byte offsets of the two source columns inside a crop row
  bool p0 = false, p1 = false;   // the column lies in the white pre-pad patch
  if (mode == 3) {
    cx = lin_coef_scaled(dx, w, g_scale_x, false);
    a0 = fix11(cx.w0);
    a1 = fix11(cx.w1);
    p0 = pre_pad && (cx.i0 < h || cx.i0 >= h + w0);
    p1 = pre_pad && (cx.i1 < h || cx.i1 >= h + w0);
    o0 = p0 ? 0 : (pre_pad ? cx.i0 - h : cx.i0) * 3;
    o1 = p1 ? 0 : (pre_pad ? cx.i1 - h : cx.i1) * 3;
  }
  // The 11-bit bilinear of cv2 is separable (resize_fix_sample): the horizontal pass of a source row, (s0 a0 + s1 a1)
  // >> 4 per channel, is kept for the next canvas rows -- consecutive rows share source rows whenever the vertical
  // scale is below ~2 (the row indices are block-uniform, so the reuse branches do not diverge).
  auto hrow = [&](int sy, int* o3) {
    const uint8_t* rowp = base + (size_t)sy * w0 * 3;
    int s0[3], s1[3];
    if (p0) { s0[0] = s0[1] = s0[2] = 255; } else { s0[0] = rowp[o0]; s0[1] = rowp[o0 + 1]; s0[2] = rowp[o0 + 2]; }
    if (p1) { s1[0] = s1[1] = s1[2] = 255; } else { s1[0] = rowp[o1]; s1[1] = rowp[o1 + 1]; s1[2] = rowp[o1 + 2]; }
#pragma unroll
    for (int k = 0; k < 3; ++k) o3[k] = (s0[k] * a0 + s1[k] * a1) >> 4;
  };
  int ya = -1, yb = -1, va[3] = {0, 0, 0}, vb[3] = {0, 0, 0};
#pragma unroll 4
  for (int r = 0; r < CANVAS_ROWS; ++r) {
    const int y = y0 + r;
    if (y >= Ht) break;
    uint8_t g = 0;
    if (mode != 0 && y >= top && y < top + rh) {
      const int dy = y - top;
      int px[3];
      if (mode == 1) {
        fetch(dy, dx, px);
      } else if (mode == 2) {
        int a[3], b[3], c[3], d[3];
        fetch(2 * dy, 2 * dx, a); fetch(2 * dy, 2 * dx + 1, b);
        fetch(2 * dy + 1, 2 * dx, c); fetch(2 * dy + 1, 2 * dx + 1, d);
        for (int k = 0; k < 3; ++k) px[k] = (a[k] + b[k] + c[k] + d[k] + 2) >> 2;
      } else {
        const int i0 = r_i0[r], i1 = r_i1[r], b0 = r_b0[r], b1 = r_b1[r];
        int h0[3], h1[3];
        if (i0 == yb) { h0[0] = vb[0]; h0[1] = vb[1]; h0[2] = vb[2]; }
        else if (i0 == ya) { h0[0] = va[0]; h0[1] = va[1]; h0[2] = va[2]; }
        else hrow(i0, h0);
        if (i1 == i0) { h1[0] = h0[0]; h1[1] = h0[1]; h1[2] = h0[2]; }
        else if (i1 == yb) { h1[0] = vb[0]; h1[1] = vb[1]; h1[2] = vb[2]; }
        else if (i1 == ya) { h1[0] = va[0]; h1[1] = va[1]; h1[2] = va[2]; }
        else hrow(i1, h1);
        ya = i0; yb = i1;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          va[k] = h0[k]; vb[k] = h1[k];
          const int v = (((b0 * h0[k]) >> 16) + ((b1 * h1[k]) >> 16) + 2) >> 2;    // = resize_fix_sample
          px[k] = min(max(v, 0), 255);
        }
      }
      g = (uint8_t)((px[0] * 9798 + px[1] * 19235 + px[2] * 3735 + 16384) >> 15);
    }
    canvas[((size_t)l * Ht + y) * Wt + x] = g;
  }
}

// ------------------------------------------------------------------------------------------
// a-9 greedy CTC: argmax (first max wins, NaN counts as max like np.argmax) + collapse + drop blank
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) {
  // true if (v, i) should replace (bv, bi): larger value, NaN beats numbers, ties -> lower index
  const bool vn = isnan(v), bn = isnan(bv);
  if (vn != bn) return vn;
  if (vn && bn) return i < bi;
  return v > bv || (v == bv && i < bi);
}

// ------------------------------------------------------------------------------------------
// Beam-search front end (BDRC/inference.py:114-118, 386-414: CTCDecoder.ctc_beam_decode): per time step the
// log-softmax of the row and its candidate tokens -- those with log p >= min_logp, at most CTC_TOPK of them, in
// descending order (ties: lower index first), the arg-max always included.  One warp per (line, step); the
// prefix merge over these few candidates per step runs on the host (ocr_inference_b200/ctc_beam.py).
// ------------------------------------------------------------------------------------------
constexpr int CTC_TOPK = 16;
__global__ void __launch_bounds__(256)
ctc_topk_kernel(const float* __restrict__ logits, int L, int V, int T, int layout, float min_logp,
                int32_t* __restrict__ tok, float* __restrict__ logp, int32_t* __restrict__ cnt) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= L * T) return;
  const int l = warp / T, t = warp - l * T;
  const float* base = logits + (size_t)l * V * T;
  const size_t sv = layout == BDRC_LAYOUT_TV ? 1 : (size_t)T;
  const float* row = base + (layout == BDRC_LAYOUT_TV ? (size_t)t * V : (size_t)t);
  // lane-strided copy of the row: v = lane, lane + 32, ...  (V <= 32 * 8)
  float x[8];
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int v = lane + 32 * j;
    x[j] = v < V ? row[(size_t)v * sv] : -INFINITY;
    m = fmaxf(m, x[j]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) sum += (lane + 32 * j < V) ? expf(x[j] - m) : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);   // xor tree: same bits in every lane
  const float lse = m + logf(sum);
  int n = 0;
  const size_t ob = ((size_t)l * T + t) * CTC_TOPK;
  for (int k = 0; k < CTC_TOPK; ++k) {
    // warp arg-max over what is left: larger value wins, on equal values the lower token index
    float bv = -INFINITY;
    int bi = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int v = lane + 32 * j;
      if (v < V && (x[j] > bv || (x[j] == bv && v < bi))) { bv = x[j]; bi = v; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (bi == 0x7fffffff) break;                       // row exhausted (or NaN / -inf only)
    const float lp = bv - lse;
    if (k > 0 && !(lp >= min_logp)) break;
    if (lane == 0) { tok[ob + k] = bi; logp[ob + k] = lp; }
    ++n;
    if ((bi & 31) == lane) x[bi >> 5] = -INFINITY;     // remove it from its owner's registers
  }
  if (lane == 0) cnt[(size_t)l * T + t] = n;
}

// [T][V] logits of a page batch (the layout the Easter2 head writes): one persistent CTA per SM walks over lines.
// A line is cut into groups of 32 time steps; warp w owns groups w and w + CTCB_WARPS and one shared-memory slot for
// each.  A group lands in its slot by ONE bulk asynchronous copy (cp.async.bulk global -> shared, completion on the
// slot's mbarrier: no registers, no instructions per byte); the warp takes the argmax of the group's rows (lane = row;
// the row pitch V is odd for the 77-symbol charset: conflict-free) and at once re-arms the slot with the same group of
// the CTA's NEXT line, so the copy of one slot runs under the argmax of the other and a whole line (154 KB) per SM
// stays in flight while the collapse of the current one runs.  Collapse: keep flags by ballot, per-group counts, a
// 16-entry prefix, kept ids written in time order.  Bit-equal to the serial scan (first maximum wins, a NaN counts as
// the maximum).  Needs 16-byte aligned groups ((V * T) % 4 == 0, aligned base), T <= 64 * CTCB_WARPS and 2 * CTCB_WARPS
// slots of 128 V bytes; bdrc_ctc_greedy falls back to the staged kernel below otherwise.
constexpr int CTCB_WARPS = 8;
constexpr int CTCB_GROUPS = 2 * CTCB_WARPS;
__global__ void __launch_bounds__(CTCB_WARPS * 32, 1)
ctc_greedy_bulk_kernel(const float* __restrict__ logits, int L, int V, int T, int blank, int32_t* __restrict__ ids,
                       int32_t* __restrict__ lens, int* errflag) {
  extern __shared__ __align__(128) uint8_t ctcb_sm[];
  tcptx::griddep_launch();   // launched as a programmatic dependent of the network's last kernel
  tcptx::griddep_wait();
  const uint32_t slot_bytes = 128u * (uint32_t)V;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ctcb_sm + (size_t)CTCB_GROUPS * slot_bytes);
  int* am = reinterpret_cast<int*>(bars + CTCB_GROUPS);         // [2][32 * CTCB_GROUPS] argmax per step
  int* wcnt = am + 2 * 32 * CTCB_GROUPS;                        // [2][CTCB_GROUPS] kept symbols per group
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = (T + 31) >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < CTCB_GROUPS; ++i) tcptx::mbar_init(tcptx::smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto fetch = [&](int l, int g) {
    const uint32_t bytes = (uint32_t)min(32, T - 32 * g) * (uint32_t)V * 4u;
    const uint32_t bar = tcptx::smem_u32(&bars[g]);
    tcptx::mbar_arrive_expect_tx(bar, bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tcptx::smem_u32(ctcb_sm + (size_t)g * slot_bytes)),
                   "l"(reinterpret_cast<uint64_t>(logits + ((size_t)l * T + 32 * g) * V)), "r"(bytes), "r"(bar)
                 : "memory");
  };
  if (lane == 0 && (int)blockIdx.x < L) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
      if (warp + h * CTCB_WARPS < G) fetch((int)blockIdx.x, warp + h * CTCB_WARPS);
  }
  uint32_t phase = 0;
  int buf = 0;
  for (int l = blockIdx.x; l < L; l += gridDim.x, buf ^= 1, phase ^= 1) {
    int* amb = am + buf * 32 * CTCB_GROUPS;
    int* wc = wcnt + buf * CTCB_GROUPS;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int g = warp + h * CTCB_WARPS;
      if (g >= G) break;
      tcptx::mbar_wait(tcptx::smem_u32(&bars[g]), phase, errflag, 9);
      if (lane < T - 32 * g) {
        const float* row = reinterpret_cast<const float*>(ctcb_sm + (size_t)g * slot_bytes) + lane * V;
        // first maximum among the numbers (strict >), first NaN beside it: a NaN wins like in np.argmax
        float bv = row[0];
        int bi = 0, ni = bv != bv ? 0 : V;
#pragma unroll 7
        for (int v = 1; v < V; ++v) {
          const float x = row[v];
          if (x > bv) { bv = x; bi = v; }                 // never true for a NaN on either side
          if (x != x) ni = min(ni, v);
        }
        amb[32 * g + lane] = ni < V ? ni : bi;
      }
      __syncwarp();
      if (lane == 0 && l + (int)gridDim.x < L) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this warp's reads of the slot before the copy's writes
        fetch(l + (int)gridDim.x, g);
      }
    }
    __syncthreads();
    uint32_t masks[2];
    int idv[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int g = warp + h * CTCB_WARPS, t = 32 * g + lane;
      const bool in = g < G && t < T;
      idv[h] = in ? amb[t] : blank;
      const bool keep = in && idv[h] != blank && (t == 0 || idv[h] != amb[t - 1]);
      masks[h] = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) wc[g] = __popc(masks[h]);
    }
    __syncthreads();
    const int cnt = lane < CTCB_GROUPS ? wc[lane] : 0;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int g = warp + h * CTCB_WARPS;
      int before = lane < g ? cnt : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
      if ((masks[h] >> lane) & 1u) ids[(size_t)l * T + before + __popc(masks[h] & ((1u << lane) - 1u))] = idv[h];
      if (g == G - 1 && lane == 0) lens[l] = before + __popc(masks[h]);
    }
    // am / wcnt are double-buffered: the next line's writes go to the other half, and a warp is two barriers
    // past this line's reads before it writes this half again
  }
}

constexpr int CTC_ROWS = 128;
__global__ void __launch_bounds__(512)
ctc_greedy_kernel(const float* __restrict__ logits, int V, int T, int layout, int blank,
                  int32_t* __restrict__ ids, int32_t* __restrict__ lens) {
  extern __shared__ int sm[];
  int* am = sm;            // [T] argmax per step
  int* scan = sm + T;      // [T] keep flags / prefix sums
  const int l = blockIdx.x;
  const float* base = logits + (size_t)l * V * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  if (layout == BDRC_LAYOUT_TV) {
    // [T][V] rows are short (V = 77 -> 308 B) and unaligned: stream the line through shared memory in
    // chunks of CTC_ROWS time steps with fully coalesced loads (every thread keeps several in flight),
    // then FOUR threads per time step scan a quarter of its row each and combine by shuffle.
    float* buf = reinterpret_cast<float*>(sm + 2 * T);
    const int pitch = V | 1;
    const int per = (V + 3) >> 2;                       // four threads share a row: V / 4 logits each
    const int step_r = 512 / V, step_c = 512 - step_r * V;   // blockDim.x == 512 (launch below)
    const int r = threadIdx.x >> 2, part = threadIdx.x & 3;
    for (int t0 = 0; t0 < T; t0 += CTC_ROWS) {
      const int rows = min(CTC_ROWS, T - t0);
      const float* src = base + (size_t)t0 * V;
      // fully coalesced flat copy, eight loads in flight per thread; (row, column) of element i advance
      // incrementally (one division per kernel instead of one per element)
      {
        const int n = rows * V;
        int rr = threadIdx.x / V, cc = threadIdx.x - rr * V;
        constexpr int BATCH = 10;   // loads in flight per thread (V = 77: two batches per 128-step chunk; 20 measured slower: 89 against 79 us)
        for (int i0 = threadIdx.x; i0 < n; i0 += 512 * BATCH) {
          float x[BATCH];
#pragma unroll
          for (int k = 0; k < BATCH; ++k) x[k] = i0 + 512 * k < n ? __ldg(src + i0 + 512 * k) : 0.f;
#pragma unroll
          for (int k = 0; k < BATCH; ++k) {
            if (i0 + 512 * k < n) buf[rr * pitch + cc] = x[k];
            rr += step_r;
            cc += step_c;
            if (cc >= V) { cc -= V; ++rr; }
          }
        }
      }
      __syncthreads();
      {
        float bv = 0.f;
        int bi = -1;                                    // -1: this thread holds no candidate
        const int v0 = part * per, v1 = min(V, v0 + per);
        if (r < rows && v0 < v1) {
          const float* row = buf + r * pitch;
          bv = row[v0];
          bi = v0;
          for (int v = v0 + 1; v < v1; ++v) {
            const float x = row[v];
            if (better(x, v, bv, bi)) { bv = x; bi = v; }
          }
        }
        // combine the four partial results of a row (lanes 4k .. 4k+3): first max wins, as in the serial scan
#pragma unroll
        for (int o = 1; o < 4; o <<= 1) {
          const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (oi >= 0 && (bi < 0 || better(ov, oi, bv, bi))) { bv = ov; bi = oi; }
        }
        if (part == 0 && r < rows) am[t0 + r] = bi;
      }
      __syncthreads();
    }
  } else {
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      float bv = base[t]; int bi = 0;
      for (int v = 1; v < V; ++v) {
        const float x = base[(size_t)v * T + t];
        if (better(x, v, bv, bi)) { bv = x; bi = v; }
      }
      am[t] = bi;
    }
  }
  __syncthreads();
  // keep[t] = id != blank && (t == 0 || id != prev); exclusive scan by one warp-per-chunk pass
  for (int t = threadIdx.x; t < T; t += blockDim.x)
    scan[t] = (am[t] != blank && (t == 0 || am[t] != am[t - 1])) ? 1 : 0;
  __syncthreads();
  if (warp == 0) {
    int carry = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
      const int t = t0 + lane;
      const int k = t < T ? scan[t] : 0;
      int incl = k;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
      }
      if (t < T && k) ids[(size_t)l * T + carry + incl - 1] = am[t];
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) lens[l] = carry;
  }
}

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int bdrc_seg_preprocess(bdrc_ctx* ctx, const uint8_t* page, int H, int W, int patch,
                        uint8_t* tiles, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  BDRC_REQUIRE(ctx, page && tiles && H > 0 && W > 0 && patch > 0 && patch % 64 == 0,
               "seg_preprocess: bad arguments (H=%d W=%d patch=%d)", H, W, patch);
  const int nty = ceil_div(H, patch), ntx = ceil_div(W, patch);
  adaptive::SrcPageTiles src{page, H, W, ntx, patch};
  adaptive::SinkTiles01 sink{tiles, patch, patch};
  bdrc_prof_begin(ctx, "a-1.adaptive_threshold", 0, (cudaStream_t)stream);
  const int rc = adaptive::launch(ctx, src, sink, nty * ntx, patch, patch, (cudaStream_t)stream);
  bdrc_prof_end(ctx, (cudaStream_t)stream);
  return rc;
}

int bdrc_seg_preprocess_batch(bdrc_ctx* ctx, const uint8_t* const* pages, const int* H, const int* W, int n_pages,
                              int patch, uint8_t* tiles, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  BDRC_REQUIRE(ctx, pages && H && W && tiles && n_pages >= 0 && patch > 0 && patch % 64 == 0,
               "seg_preprocess_batch: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  size_t tile_off = 0;
  for (int p0 = 0; p0 < n_pages; p0 += adaptive::MAX_BATCH_PAGES) {
    const int n = n_pages - p0 < adaptive::MAX_BATCH_PAGES ? n_pages - p0 : adaptive::MAX_BATCH_PAGES;
    adaptive::SrcPageTable src;
    memset(&src, 0, sizeof(src));
    src.n_pages = n;
    src.patch = patch;
    int t = 0;
    for (int k = 0; k < n; ++k) {
      BDRC_REQUIRE(ctx, pages[p0 + k] && H[p0 + k] > 0 && W[p0 + k] > 0, "seg_preprocess_batch: page %d is empty", p0 + k);
      src.page[k] = pages[p0 + k];
      src.Hc[k] = H[p0 + k];
      src.Wc[k] = W[p0 + k];
      src.ntx[k] = ceil_div(W[p0 + k], patch);
      src.tile0[k] = t;
      t += ceil_div(H[p0 + k], patch) * src.ntx[k];
    }
    src.tile0[n] = t;
    adaptive::SinkTiles01 sink{tiles + tile_off * (size_t)patch * patch, patch, patch};
    bdrc_prof_begin(ctx, "a-1.adaptive_threshold", 0, s);
    const int rc = adaptive::launch(ctx, src, sink, t, patch, patch, s);
    bdrc_prof_end(ctx, s);
    if (rc) return rc;
    tile_off += (size_t)t;
  }
  return 0;
}

int bdrc_seg_tile_rgb(bdrc_ctx* ctx, const uint8_t* page, int H, int W, int Hc, int Wc, int patch,
                      float* tiles, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  BDRC_REQUIRE(ctx, page && tiles && H > 0 && W > 0 && Hc > 0 && Wc > 0 && patch > 0,
               "seg_tile_rgb: bad arguments");
  const int nty = ceil_div(Hc, patch), ntx = ceil_div(Wc, patch);
  dim3 grid(ceil_div(patch, 128), patch, nty * ntx);
  tile_rgb_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(page, H, W, Hc, Wc, patch, ntx, tiles);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

static int seg_post_common(bdrc_ctx* ctx, const float* logits, int nty, int ntx, int C, int patch,
                           int lowres, int Hc, int Wc, int H0, int W0, int mode, float thr,
                           uint8_t* mask, int torch_path, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  BDRC_REQUIRE(ctx, logits && mask && C >= 1 && C <= 32 && nty > 0 && ntx > 0,
               "seg_postprocess: bad arguments (C=%d)", C);
  BDRC_REQUIRE(ctx, Hc <= nty * patch && Wc <= ntx * patch && Hc > (nty - 1) * patch &&
                        Wc > (ntx - 1) * patch && H0 > 0 && W0 > 0,
               "seg_postprocess: cropped size %dx%d inconsistent with %dx%d tiles", Hc, Wc, nty, ntx);
  PostParams p;
  p.logits = logits; p.ntx = ntx; p.C = C; p.P = patch; p.lowres = lowres;
  p.Hc = Hc; p.Wc = Wc; p.H0 = H0; p.W0 = W0; p.mode = mode; p.thr = thr;
  p.thr_logit = sigmoid_logit_threshold(thr);
  p.mask = mask;
  cudaStream_t s = (cudaStream_t)stream;
  bdrc_prof_begin(ctx, "a-3.seg_post", 0, s);
  if (torch_path) {
    dim3 grid(ceil_div(W0, 128), H0);
    seg_post_torch_kernel<<<grid, 128, 0, s>>>(p);
  } else if (mode == BDRC_POST_SIGMOID && C == 1 && lowres && Hc == H0 && Wc == W0 && patch % 16 == 0) {
    dim3 grid(ceil_div(ceil_div(W0, 16), 128), H0);
    seg_post_line_lowres_kernel<<<grid, 128, 0, s>>>(p);
  } else if (mode == BDRC_POST_SIGMOID && C == 1 && !lowres && Hc == H0 && Wc == W0 &&
             patch % 4 == 0) {
    dim3 grid(ceil_div(ceil_div(W0, 4), 128), H0);
    seg_post_line_fast_kernel<<<grid, 128, 0, s>>>(logits, ntx, patch, H0, W0, p.thr_logit, mask);
  } else {
    dim3 grid(ceil_div(W0, 128), H0);
    seg_post_kernel<<<grid, 128, 0, s>>>(p);
  }
  bdrc_prof_end(ctx, s);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

int bdrc_seg_postprocess(bdrc_ctx* ctx, const float* logits, int nty, int ntx, int C, int patch,
                         int Hc, int Wc, int H0, int W0, int mode, float thr, uint8_t* mask,
                         void* stream) {
  return seg_post_common(ctx, logits, nty, ntx, C, patch, 0, Hc, Wc, H0, W0, mode, thr, mask, 0, stream);
}

int bdrc_seg_postprocess_lowres(bdrc_ctx* ctx, const float* logits_lowres, int nty, int ntx, int C,
                                int patch, int Hc, int Wc, int mode, float thr, uint8_t* mask,
                                void* stream) {
  return seg_post_common(ctx, logits_lowres, nty, ntx, C, patch, 1, Hc, Wc, Hc, Wc, mode, thr, mask, 0,
                         stream);
}

int bdrc_seg_postprocess_lowres_batch(bdrc_ctx* ctx, const float* const* logits_lowres, const int* nty,
                                      const int* ntx, const int* Hc, const int* Wc, int n_pages, int C, int patch,
                                      int mode, float thr, uint8_t* const* masks, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  BDRC_REQUIRE(ctx, logits_lowres && nty && ntx && Hc && Wc && masks && n_pages >= 0, "seg_postprocess_lowres_batch: bad arguments");
  if (!(mode == BDRC_POST_SIGMOID && C == 1 && patch % 16 == 0)) {
    // layout models (several classes / softmax): page by page through the generic kernel
    for (int k = 0; k < n_pages; ++k)
      BDRC_TRY(seg_post_common(ctx, logits_lowres[k], nty[k], ntx[k], C, patch, 1, Hc[k], Wc[k], Hc[k], Wc[k], mode, thr,
                               masks[k], 0, stream));
    return 0;
  }
  cudaStream_t s = (cudaStream_t)stream;
  for (int p0 = 0; p0 < n_pages; p0 += POST_MAX_PAGES) {
    const int n = n_pages - p0 < POST_MAX_PAGES ? n_pages - p0 : POST_MAX_PAGES;
    PostBatch b;
    memset(&b, 0, sizeof(b));
    b.P = patch;
    b.thr_logit = sigmoid_logit_threshold(thr);
    int max_w = 0, max_h = 0;
    for (int k = 0; k < n; ++k) {
      const int i = p0 + k;
      BDRC_REQUIRE(ctx, logits_lowres[i] && masks[i] && nty[i] > 0 && ntx[i] > 0 && Hc[i] <= nty[i] * patch &&
                            Wc[i] <= ntx[i] * patch && Hc[i] > (nty[i] - 1) * patch && Wc[i] > (ntx[i] - 1) * patch,
                   "seg_postprocess_lowres_batch: page %d: size %dx%d inconsistent with %dx%d tiles", i, Hc[i], Wc[i],
                   nty[i], ntx[i]);
      b.pg[k] = PostPage{logits_lowres[i], masks[i], ntx[i], Hc[i], Wc[i]};
      if (Wc[i] > max_w) max_w = Wc[i];
      if (Hc[i] > max_h) max_h = Hc[i];
    }
    dim3 grid((unsigned)ceil_div(ceil_div(max_w, 16), 64), (unsigned)ceil_div(max_h, 16), (unsigned)n);
    bdrc_prof_begin(ctx, "a-3.seg_post", 0, s);
    seg_post_line_lowres_batch_kernel<<<grid, dim3(64, 4), 0, s>>>(b);
    bdrc_prof_end(ctx, s);
    BDRC_CHECK_LAUNCH(ctx);
  }
  return 0;
}

int bdrc_seg_postprocess_torch(bdrc_ctx* ctx, const float* logits, int nty, int ntx, int C, int patch,
                               int Hc, int Wc, int H0, int W0, float thr, uint8_t* mask,
                               void* stream) {
  return seg_post_common(ctx, logits, nty, ntx, C, patch, 0, Hc, Wc, H0, W0, BDRC_POST_SOFTMAX, thr,
                         mask, 1, stream);
}

int bdrc_resize_bilinear_u8(bdrc_ctx* ctx, const uint8_t* src, int Hs, int Ws, int C, uint8_t* dst,
                            int Hd, int Wd, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  BDRC_REQUIRE(ctx, src && dst && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0 && C > 0 && C <= 8,
               "resize_bilinear_u8: bad arguments");
  dim3 grid(ceil_div(Wd, 128), Hd);
  resize_u8_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(src, Hs, Ws, C, dst, Hd, Wd);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

int bdrc_resize_bilinear_f32(bdrc_ctx* ctx, const uint8_t* src, int Hs, int Ws, int C, float* dst,
                             int Hd, int Wd, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  BDRC_REQUIRE(ctx, src && dst && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0 && C > 0,
               "resize_bilinear_f32: bad arguments");
  dim3 grid(ceil_div(Wd, 128), Hd);
  resize_f32_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(src, Hs, Ws, C, dst, Hd, Wd);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

size_t bdrc_ocr_prepare_workspace_bytes(int L, int Ht, int Wt) {
  return align_up((size_t)L * Ht * Wt, 256) + align_up((size_t)L * sizeof(int), 256);
}

int bdrc_ocr_prepare_lines(bdrc_ctx* ctx, const uint8_t* crops, const int32_t* crop_hw,
                           const int64_t* crop_off, int L, int Ht, int Wt, int pre_pad, int Cpad,
                           void* lines_bf16, float* lines_f32, void* workspace,
                           size_t workspace_bytes, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  if (L == 0) return 0;
  BDRC_REQUIRE(ctx, crops && crop_hw && crop_off && L > 0 && Ht > 0 && Wt > 0 && (Cpad >= Ht || !lines_bf16) &&
                        (lines_bf16 || lines_f32),
               "ocr_prepare_lines: bad arguments");
  BDRC_REQUIRE(ctx, workspace && workspace_bytes >= bdrc_ocr_prepare_workspace_bytes(L, Ht, Wt),
               "ocr_prepare_lines: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  uint8_t* canvas = reinterpret_cast<uint8_t*>(workspace);
  int* status = reinterpret_cast<int*>(canvas + align_up((size_t)L * Ht * Wt, 256));
  // status[l] is written by the canvas kernel for every line: no memset node between the kernels of consecutive calls.
  // Programmatic dependent launch of canvas / threshold / CTC was measured (every kernel waits for its predecessor
  // before it touches memory): 0.7062 against 0.7063 ms per 64-line step, no gain -- and early-resident CTAs of
  // these kernels could hold SM slots the segmentation convs of the page pipeline want; left off (numAttrs = 0).
  cudaLaunchAttribute pdl_attr[1];
  pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ceil_div(Wt, 128), ceil_div(Ht, CANVAS_ROWS), L);
  cfg.blockDim = dim3(128);
  cfg.stream = s;
  cfg.attrs = pdl_attr;
  cfg.numAttrs = 0;
  bdrc_prof_begin(ctx, "a-7.canvas", 0, s);
  BDRC_CUDA(ctx, cudaLaunchKernelEx(&cfg, ocr_canvas_kernel, crops, crop_hw, crop_off, Ht, Wt, pre_pad, canvas, status));
  bdrc_prof_end(ctx, s);
  BDRC_CHECK_LAUNCH(ctx);
  // (the padding channels Ht..Cpad-1 of lines_bf16 are zeroed by the threshold kernel's sink, which covers every
  //  column of every line)
  adaptive::SrcGrayBatch src{canvas, Ht, Wt};
  adaptive::SinkOcrLines sink{reinterpret_cast<bf16*>(lines_bf16), lines_f32, Ht, Wt, Cpad};
  bdrc_prof_begin(ctx, "a-7.adaptive_threshold", 0, s);
  const int rc = adaptive::launch(ctx, src, sink, L, Ht, Wt, s, /*pdl=*/false);
  bdrc_prof_end(ctx, s);
  return rc;
}

int bdrc_ctc_greedy(bdrc_ctx* ctx, const float* logits, int L, int V, int T, int layout, int blank,
                    int32_t* ids, int32_t* lens, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  if (L == 0) return 0;
  BDRC_REQUIRE(ctx, logits && ids && lens && V > 0 && T > 0 && T <= 20000 &&
                        (layout == BDRC_LAYOUT_TV || layout == BDRC_LAYOUT_VT),
               "ctc_greedy: bad arguments (V=%d T=%d)", V, T);
  if (layout == BDRC_LAYOUT_TV && T <= 32 * CTCB_GROUPS && ((long long)V * T) % 4 == 0 &&
      (reinterpret_cast<uintptr_t>(logits) & 15) == 0) {
    const size_t bsmem = (size_t)CTCB_GROUPS * 128 * V + CTCB_GROUPS * sizeof(uint64_t) +
                         (size_t)(2 * 32 * CTCB_GROUPS + 2 * CTCB_GROUPS) * sizeof(int);
    if (bsmem <= 220 * 1024) {
      static size_t attr_bytes = 0;
      if (bsmem > attr_bytes) {
        BDRC_CUDA(ctx, cudaFuncSetAttribute(ctc_greedy_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr_bytes = 220 * 1024;
      }
      bdrc_prof_begin(ctx, "a-9.ctc_greedy", 0, (cudaStream_t)stream);
      cudaLaunchAttribute pdl_attr[1];
      pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(L < ctx->sm_count ? L : ctx->sm_count));
      cfg.blockDim = dim3(CTCB_WARPS * 32);
      cfg.dynamicSmemBytes = bsmem;
      cfg.stream = (cudaStream_t)stream;
      cfg.attrs = pdl_attr;
      cfg.numAttrs = 0;   // see bdrc_ocr_prepare_lines
      BDRC_CUDA(ctx, cudaLaunchKernelEx(&cfg, ctc_greedy_bulk_kernel, logits, L, V, T, blank, ids, lens, ctx->d_errflag));
      bdrc_prof_end(ctx, (cudaStream_t)stream);
      BDRC_CHECK_LAUNCH(ctx);
      return 0;
    }
  }
  const size_t smem = (size_t)2 * T * sizeof(int) +
                      (layout == BDRC_LAYOUT_TV ? (size_t)CTC_ROWS * (V | 1) * sizeof(float) : 0);
  BDRC_REQUIRE(ctx, smem <= 200 * 1024, "ctc_greedy: V * 128 + 2 T exceeds the shared-memory budget (V=%d T=%d)", V, T);
  if (smem > 48 * 1024)
    BDRC_CUDA(ctx, cudaFuncSetAttribute(ctc_greedy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
  bdrc_prof_begin(ctx, "a-9.ctc_greedy", 0, (cudaStream_t)stream);
  ctc_greedy_kernel<<<L, 512, smem, (cudaStream_t)stream>>>(logits, V, T, layout, blank, ids, lens);
  bdrc_prof_end(ctx, (cudaStream_t)stream);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

int bdrc_ctc_topk(bdrc_ctx* ctx, const float* logits, int L, int V, int T, int layout, float min_logp,
                  int32_t* tokens, float* logps, int32_t* counts, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  if (L == 0) return 0;
  BDRC_REQUIRE(ctx, logits && tokens && logps && counts && V > 0 && V <= 256 && T > 0 &&
                        (layout == BDRC_LAYOUT_TV || layout == BDRC_LAYOUT_VT),
               "ctc_topk: bad arguments (V=%d T=%d; V <= 256)", V, T);
  const long long warps = (long long)L * T;
  bdrc_prof_begin(ctx, "a-9.ctc_topk", 0, (cudaStream_t)stream);
  ctc_topk_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>(logits, L, V, T, layout, min_logp, tokens,
                                                                                 logps, counts);
  bdrc_prof_end(ctx, (cudaStream_t)stream);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

int bdrc_ctc_topk_width(void) { return CTC_TOPK; }

}  // extern "C"
