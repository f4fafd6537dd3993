// cv2.adaptiveThreshold(ADAPTIVE_THRESH_GAUSSIAN_C, THRESH_BINARY, block=51, C=13) as one fused
// kernel (BDRC/utils.py:796-815): gray source -> fp32 separable 51-tap Gaussian with replicate
// border at the IMAGE (= tile) edge -> mean rounded half-to-even -> gray - mean > -13.
//
// Arithmetic is pinned to oracle/hostops.py:adaptive_threshold: row pass k = 0..50 sequential
// FMA, column pass centre tap then symmetric pairs (x[-k] + x[k]) * w[k] by FMA.
//
// Work decomposition: one CTA owns a 32-column strip of one image and STREAMS down its rows in
// blocks of 32.  Per block the gray halo (32 rows x 82 columns, replicate-clamped at the image edge)
// is staged as fp32, the row pass writes 32 rows of a 128-row shared-memory ring, and -- two blocks
// behind, when its 51-row window is complete -- the column pass + threshold emits 32 output rows.
// Nothing is recomputed vertically (the first version recomputed the row pass for a 114-row halo of
// every 64-row region: 142 flop per pixel; now 102), both passes keep sliding register windows
// (~1 shared load per 4 FMAs) and ~40 KB of shared memory allows 4 CTAs per SM.
// The kernel is fp32-FMA bound, not HBM bound: 102 flop per pixel against 4 bytes of DRAM traffic.
// Round 2 measured a packed variant (fma/add/mul.rn.f32x2 on row pairs in the row pass and column pairs in the column
// pass, bit-identical results, 28 % fewer instructions): 1.27 ms against 1.12 ms per 16 pages.  FFMA2 / FADD2 occupy
// the FMA pipe twice as long as the scalar forms, and that pipe, not instruction issue, is what this kernel saturates
// (102 FMA-pipe operations per pixel + integer multiply-adds of the addressing on the same pipe).
#pragma once
#include <cmath>

#include "common.cuh"

namespace adaptive {

constexpr int KS = 51, KR = 25;
constexpr int SW = 32;                 // strip width (output columns per CTA)
constexpr int RB = 32;                 // rows per block
constexpr int GW = SW + 2 * KR;        // 82 staged columns
constexpr int GSTRIDE = 92;            // padded row stride of the staged gray block (float4-aligned)
constexpr int GOFF = 4;                // staged column c lives at g[row * GSTRIDE + GOFF + c]: a quad store that hangs over the
                                       // 82 columns by up to 3 on either side lands in padding, no predicates
constexpr int RING = 128;              // row-pass ring (rows), power of two
constexpr int WIN = 2 * KR + 4;        // ring rows under the four output rows of a column-pass thread
constexpr int RING_PHYS = RING + WIN;  // rows of the buffer: slots 0 .. WIN-1 are kept twice (at s and at s + RING), so a
                                       // window that starts anywhere in the ring is WIN consecutive rows of the buffer
constexpr int THREADS = 256;
constexpr size_t SMEM_BYTES = (size_t)(RB * GSTRIDE + RING_PHYS * SW) * sizeof(float) + RING * SW + RB * SW;

__constant__ float c_gauss[KS];

// Source policies ------------------------------------------------------------------------------
// A policy hands out a per-CTA View of image `img` (one tile / one line): everything that depends on the image
// index -- the page of a multi-page batch, the tile's origin in it -- is resolved once per CTA, not per pixel.
//
// RGB page viewed as row-major 512x512 tiles, padded with 255 beyond (Hc, Wc) (utils.py:775-776)
struct PageTileView {
  const uint8_t* page;
  int Hc, Wc, py0, px0;
  __device__ __forceinline__ float gray(int y, int x) const {
    const int py = py0 + y, px = px0 + x;
    if (py >= Hc || px >= Wc) return 255.f;
    const uint8_t* p = page + ((size_t)py * Wc + px) * 3;
    const int g = (p[0] * 9798 + p[1] * 19235 + p[2] * 3735 + 16384) >> 15;
    return (float)g;
  }
  // four horizontally adjacent pixels x .. x+3 of one tile row (x % 4 == 0, all inside the tile) as raw words:
  // three 32-bit loads instead of twelve byte loads.  Returns 1 (raw holds the 12 RGB bytes), 2 (all four
  // pixels lie in the 255 padding) or 0 where the caller must go pixel by pixel (quad straddling the page
  // edge, or a row that is not 4-byte aligned).  conv4 turns the raw words into gray values.
  static constexpr int WORDS = 3;
  __device__ __forceinline__ int load4(int y, int x, uint32_t* raw) const {
    const int py = py0 + y, px = px0 + x;
    if (py >= Hc || px >= Wc) return 2;
    const uint8_t* p = page + ((size_t)py * Wc + px) * 3;
    if (px + 3 >= Wc || (reinterpret_cast<uintptr_t>(p) & 3)) return 0;
    const uint32_t* q = reinterpret_cast<const uint32_t*>(p);
    raw[0] = __ldg(q); raw[1] = __ldg(q + 1); raw[2] = __ldg(q + 2);
    return 1;
  }
  __device__ __forceinline__ void conv4(const uint32_t* raw, float* v, uint32_t& packed) const {
    const uint32_t w0 = raw[0], w1 = raw[1], w2 = raw[2];
    // bytes: w0 = r0 g0 b0 r1, w1 = g1 b1 r2 g2, w2 = b2 r3 g3 b3 (little endian).  One register per pixel with
    // (r, g, b, x) in its bytes, then the 15-bit gray as two byte dot products: 9798 = 38*256 + 70,
    // 19235 = 75*256 + 35, 3735 = 14*256 + 151 (the fourth byte meets a zero coefficient)
    const uint32_t p0 = w0, p1 = __funnelshift_r(w0, w1, 24), p2 = __funnelshift_r(w1, w2, 16), p3 = w2 >> 8;
    constexpr uint32_t CH = 38u | (75u << 8) | (14u << 16), CL = 70u | (35u << 8) | (151u << 16);
    const uint32_t g0 = ((__dp4a(p0, CH, 0u) << 8) + __dp4a(p0, CL, 16384u)) >> 15;
    const uint32_t g1 = ((__dp4a(p1, CH, 0u) << 8) + __dp4a(p1, CL, 16384u)) >> 15;
    const uint32_t g2 = ((__dp4a(p2, CH, 0u) << 8) + __dp4a(p2, CL, 16384u)) >> 15;
    const uint32_t g3 = ((__dp4a(p3, CH, 0u) << 8) + __dp4a(p3, CL, 16384u)) >> 15;
    v[0] = (float)g0; v[1] = (float)g1; v[2] = (float)g2; v[3] = (float)g3;
    packed = g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
  }
};

struct SrcPageTiles {
  const uint8_t* page;
  int Hc, Wc, ntx, patch;
  typedef PageTileView View;
  int ih() const { return patch; }
  int iw() const { return patch; }
  __device__ __forceinline__ View view(int img) const {
    const int ty = img / ntx, tx = img - ty * ntx;
    return View{page, Hc, Wc, ty * patch, tx * patch};
  }
};

// Several pages in ONE launch (bdrc_seg_preprocess_batch): tile `img` of the batch belongs to the page whose
// tile range [tile0[k], tile0[k + 1]) holds it.  The table travels in the kernel parameters.
constexpr int MAX_BATCH_PAGES = 32;
struct SrcPageTable {
  const uint8_t* page[MAX_BATCH_PAGES];
  int Hc[MAX_BATCH_PAGES], Wc[MAX_BATCH_PAGES], ntx[MAX_BATCH_PAGES], tile0[MAX_BATCH_PAGES + 1];
  int n_pages, patch;
  typedef PageTileView View;
  int ih() const { return patch; }
  int iw() const { return patch; }
  __device__ __forceinline__ View view(int img) const {
    int k = 0;
    while (k + 1 < n_pages && img >= tile0[k + 1]) ++k;
    const int t = img - tile0[k];
    const int ty = t / ntx[k], tx = t - ty * ntx[k];
    return View{page[k], Hc[k], Wc[k], ty * patch, tx * patch};
  }
};

// batch of single-channel u8 images [n][H][W]
struct GrayView {
  const uint8_t* img;
  int W;
  __device__ __forceinline__ float gray(int y, int x) const { return (float)img[(size_t)y * W + x]; }
  static constexpr int WORDS = 1;
  __device__ __forceinline__ int load4(int y, int x, uint32_t* raw) const {
    const uint8_t* p = img + (size_t)y * W + x;
    if (reinterpret_cast<uintptr_t>(p) & 3) return 0;
    raw[0] = __ldg(reinterpret_cast<const uint32_t*>(p));
    return 1;
  }
  __device__ __forceinline__ void conv4(const uint32_t* raw, float* v, uint32_t& packed) const {
    const uint32_t w = raw[0];
    v[0] = (float)(w & 255u); v[1] = (float)((w >> 8) & 255u); v[2] = (float)((w >> 16) & 255u); v[3] = (float)(w >> 24);
    packed = w;
  }
};

struct SrcGrayBatch {
  const uint8_t* img;
  int H, W;
  typedef GrayView View;
  int ih() const { return H; }
  int iw() const { return W; }
  __device__ __forceinline__ View view(int i) const { return View{img + (size_t)i * H * W, W}; }
};

// Sink policies --------------------------------------------------------------------------------
// `res` is a block of RB rows x SW columns of bytes (1 = white) whose top-left output pixel is (y0, x0).
// u8 {0,1} tiles [n][H][W]
struct SinkTiles01 {
  uint8_t* out;
  int H, W;
  __device__ __forceinline__ void store(const uint8_t* res, int img, int y0, int x0, int IH, int IW) const {
    for (int i = threadIdx.x; i < RB * SW / 4; i += THREADS) {
      const int r = i / (SW / 4), c4 = (i % (SW / 4)) * 4;
      const int y = y0 + r, x = x0 + c4;
      if (y >= IH || x >= IW) continue;
      uint8_t* o = out + ((size_t)img * H + y) * W + x;
      if (x + 3 < IW && (W & 3) == 0) {
        *reinterpret_cast<uint32_t*>(o) = *reinterpret_cast<const uint32_t*>(res + r * SW + c4);
      } else {
        for (int j = 0; j < 4 && x + j < IW; ++j) o[j] = res[r * SW + c4 + j];
      }
    }
  }
};

// OCR network input: bf16 time-major [n][W][Cpad] with value 2*b-1 (inference.py:343-345) and,
// optionally, the reference layout fp32 [n][H][W].
struct SinkOcrLines {
  bf16* lines;      // may be null
  float* lines_f32; // may be null
  int H, W, Cpad;
  __device__ __forceinline__ void store(const uint8_t* res, int img, int y0, int x0, int IH, int IW) const {
    if (lines) {
      // lanes run along y (the channel axis of the output) so stores are contiguous
      for (int i = threadIdx.x; i < RB * SW; i += THREADS) {
        const int c = i / RB, r = i % RB;   // c = x offset, r = y offset
        const int y = y0 + r, x = x0 + c;
        if (y >= IH || x >= IW) continue;
        lines[((size_t)img * W + x) * Cpad + y] = __float2bfloat16(res[r * SW + c] ? 1.f : -1.f);
      }
      // the padding channels IH .. Cpad-1 of this strip's columns, once, by the CTA's last block (a memset of the
      // whole [n][W][Cpad] buffer did this before: 33 MB written twice per 64 lines)
      if (y0 + RB >= IH && Cpad > IH) {
        const int npad = Cpad - IH;
        if (((IH | Cpad) & 7) == 0 && (reinterpret_cast<uintptr_t>(lines) & 15) == 0) {          // 16-byte stores: (x * Cpad + IH) * 2 B is 16-byte aligned
          const int q = npad / 8;
          for (int i = threadIdx.x; i < SW * q; i += THREADS) {
            const int c = i / q, k = i - c * q;
            const int x = x0 + c;
            if (x < IW)
              *reinterpret_cast<uint4*>(lines + ((size_t)img * W + x) * Cpad + IH + 8 * k) = make_uint4(0u, 0u, 0u, 0u);
          }
        } else {
          for (int i = threadIdx.x; i < SW * npad; i += THREADS) {
            const int c = i / npad, k = i - c * npad;
            const int x = x0 + c;
            if (x < IW) lines[((size_t)img * W + x) * Cpad + IH + k] = __float2bfloat16(0.f);
          }
        }
      }
    }
    if (lines_f32) {
      for (int i = threadIdx.x; i < RB * SW; i += THREADS) {
        const int r = i / SW, c = i % SW;
        const int y = y0 + r, x = x0 + c;
        if (y >= IH || x >= IW) continue;
        lines_f32[((size_t)img * H + y) * W + x] = res[r * SW + c] ? 1.f : -1.f;
      }
    }
  }
};

template <class Src, class Sink>
__global__ void __launch_bounds__(THREADS, 3)
adaptive_threshold_kernel(const Src src, const Sink sink, int IH, int IW, int strips_x) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  // programmatic dependent launch (launch(..., pdl = true): the OCR line chain); no-ops in a plain launch
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  float* g = reinterpret_cast<float*>(smem_raw);                 // [RB][GSTRIDE]  staged gray block
  float* ring = g + RB * GSTRIDE;                                // [RING_PHYS][SW] row-pass results
  uint8_t* gc = reinterpret_cast<uint8_t*>(ring + RING_PHYS * SW);   // [RING][SW]  gray of the strip itself
  uint8_t* res = gc + RING * SW;                                 // [RB][SW]

  const int sx = blockIdx.x % strips_x;
  const int img = blockIdx.x / strips_x;
  const int x0 = sx * SW;
  const typename Src::View sv = src.view(img);
  const int n_blocks = (IH + RB - 1) / RB;
  // Where the row-pass result of image row y lives.  Large images: ring slot y % RING (+ the copy at slot + RING for
  // slots < WIN), windows clamped at the image border take the slow path of the column pass.  Small images (an
  // 80-row OCR canvas: four of its five column-pass warps sit at a border): row y at buffer row y + KR, the KR rows
  // above / below it filled with copies of row 0 / row IH-1 when those are computed -- the replicate border
  // materialised, every window WIN consecutive buffer rows at immediate offsets.
  const bool small = IH + 2 * KR <= RING_PHYS;

  // Staging works on quads: tile columns x0 - 28 + 4q .. + 3 (a 4-aligned superset of the 82 staged columns),
  // staged column c = 4q - 3 + j; quads 7 .. 14 are the strip itself.  The raw words of block b + 1 are loaded
  // into registers BEFORE the filter passes of block b and converted / stored after them, so the global-load
  // latency (the top stall of the previous version: long scoreboard) hides behind ~100 FMAs per pixel.
  constexpr int NQ = (GW + 3 + 3) / 4;                       // 22
  constexpr int ITEMS = RB * NQ;                             // 704
  constexpr int NIT = (ITEMS + THREADS - 1) / THREADS;       // 3
  uint32_t raw[NIT][Src::View::WORDS];
  uint32_t mode = 0;                                         // 2 bits per item: 0 slow, 1 raw words, 2 constant 255
  auto prefetch = [&](int blk) {
    mode = 0;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = threadIdx.x + it * THREADS;
      if (i < ITEMS) {
        const int r = i / NQ, q = i - r * NQ;
        const int y = min(RB * blk + r, IH - 1);
        const int xq = x0 - 28 + 4 * q;
        if (xq >= 0 && xq + 3 < IW) mode |= (uint32_t)sv.load4(y, xq, raw[it]) << (2 * it);
      }
    }
  };
  auto commit = [&](int blk) {
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int i = threadIdx.x + it * THREADS;
      if (i < ITEMS) {
        const int r = i / NQ, q = i - r * NQ;
        const int xq = x0 - 28 + 4 * q;
        const uint32_t m = (mode >> (2 * it)) & 3u;
        float v[4];
        uint32_t packed;
        if (m == 1u) {
          sv.conv4(raw[it], v, packed);
        } else if (m == 2u) {
          v[0] = v[1] = v[2] = v[3] = 255.f;
          packed = 0xFFFFFFFFu;
        } else {
          const int y = min(RB * blk + r, IH - 1);
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = sv.gray(y, min(max(xq + j, 0), IW - 1));
          packed = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
        }
        float* gr = g + r * GSTRIDE + GOFF + 4 * q - 3;      // 4q + 1 .. 4q + 4 of a 92-float row: always inside
#pragma unroll
        for (int j = 0; j < 4; ++j) gr[j] = v[j];
        if (q >= 7 && q < 7 + SW / 4)
          *reinterpret_cast<uint32_t*>(gc + ((RB * blk + r) & (RING - 1)) * SW + 4 * (q - 7)) = packed;
      }
    }
  };
  if (n_blocks > 0) prefetch(0);

  for (int b = 0; b < n_blocks + 2; ++b) {
    // ---- stage gray rows [RB*b, RB*b + RB) x columns [x0 - KR, x0 + SW + KR), clamped to the image ----
    if (b < n_blocks) commit(b);
    __syncthreads();
    if (b + 1 < n_blocks) prefetch(b + 1);

    // ---- row pass of block b: ring[row][c] = sum_k g[row][c + k] * w[k], k ascending, FMA ----
    // (rows past the image bottom -- 16 of the 96 streamed rows of an 80-row OCR canvas -- are skipped by whole
    //  warps: nobody reads their ring rows, the column pass clamps to IH - 1)
    if (b < n_blocks && RB * b + (int)(threadIdx.x / (SW / 4)) < IH) {
      const int r = threadIdx.x / (SW / 4), c = (threadIdx.x % (SW / 4)) * 4;
      const float* gp = g + r * GSTRIDE + GOFF + c;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      float4 v = *reinterpret_cast<const float4*>(gp);
      float w0 = v.x, w1 = v.y, w2 = v.z, w3 = v.w;   // sliding window g[c+k .. c+k+3]
#pragma unroll
      for (int k4 = 0; k4 < 52; k4 += 4) {
        const float4 n = *reinterpret_cast<const float4*>(gp + k4 + 4);
        // taps k4 .. k4+3 (k = 51 does not exist: guarded at compile time)
        {
          const float wk = c_gauss[k4];
          a0 = __fmaf_rn(w0, wk, a0); a1 = __fmaf_rn(w1, wk, a1);
          a2 = __fmaf_rn(w2, wk, a2); a3 = __fmaf_rn(w3, wk, a3);
        }
        if (k4 + 1 < KS) {
          const float wk = c_gauss[k4 + 1];
          a0 = __fmaf_rn(w1, wk, a0); a1 = __fmaf_rn(w2, wk, a1);
          a2 = __fmaf_rn(w3, wk, a2); a3 = __fmaf_rn(n.x, wk, a3);
        }
        if (k4 + 2 < KS) {
          const float wk = c_gauss[k4 + 2];
          a0 = __fmaf_rn(w2, wk, a0); a1 = __fmaf_rn(w3, wk, a1);
          a2 = __fmaf_rn(n.x, wk, a2); a3 = __fmaf_rn(n.y, wk, a3);
        }
        if (k4 + 3 < KS) {
          const float wk = c_gauss[k4 + 3];
          a0 = __fmaf_rn(w3, wk, a0); a1 = __fmaf_rn(n.x, wk, a1);
          a2 = __fmaf_rn(n.y, wk, a2); a3 = __fmaf_rn(n.z, wk, a3);
        }
        w0 = n.x; w1 = n.y; w2 = n.z; w3 = n.w;
      }
      const float4 acc = make_float4(a0, a1, a2, a3);
      const int y = RB * b + r;
      if (small) {
        *reinterpret_cast<float4*>(ring + (y + KR) * SW + c) = acc;
        if (y == 0) {
#pragma unroll
          for (int k = 0; k < KR; ++k) *reinterpret_cast<float4*>(ring + k * SW + c) = acc;
        }
        if (y == IH - 1) {
#pragma unroll
          for (int k = 1; k <= KR; ++k) *reinterpret_cast<float4*>(ring + (y + KR + k) * SW + c) = acc;
        }
      } else {
        const int sl = y & (RING - 1);
        *reinterpret_cast<float4*>(ring + sl * SW + c) = acc;
        if (sl < WIN) *reinterpret_cast<float4*>(ring + (sl + RING) * SW + c) = acc;
      }
    }

    // ---- column pass + threshold of block b - 2 (its 51-row window was complete one barrier ago) ----
    const int ob = b - 2;
    if (ob >= 0 && ob < n_blocks && RB * ob + (int)(threadIdx.x / SW) * 4 < IH) {   // warp-uniform; the sinks never read res rows >= IH
      const int c = threadIdx.x % SW, r4 = RB * ob + (threadIdx.x / SW) * 4;   // output rows r4 .. r4+3
      float v[WIN];
      if (small || (r4 - KR >= 0 && r4 + 3 + KR < IH)) {
        // warp-uniform (one r4 per warp); WIN consecutive buffer rows: immediate offsets, no index math
        const float* rp = ring + (small ? r4 : ((r4 - KR) & (RING - 1))) * SW + c;
#pragma unroll
        for (int i = 0; i < WIN; ++i) v[i] = rp[i * SW];
      } else {   // replicate border at the top / bottom of a large image
#pragma unroll
        for (int i = 0; i < 2 * KR + 4; ++i) v[i] = ring[(min(max(r4 - KR + i, 0), IH - 1) & (RING - 1)) * SW + c];
      }
      float s[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) s[j] = __fmul_rn(v[j + KR], c_gauss[KR]);
#pragma unroll
      for (int k = 1; k <= KR; ++k) {
        const float wk = c_gauss[KR + k];
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] = __fmaf_rn(__fadd_rn(v[j + KR - k], v[j + KR + k]), wk, s[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int mean = __float2int_rn(s[j]);
        mean = min(max(mean, 0), 255);
        const int gv = (int)gc[((r4 + j) & (RING - 1)) * SW + c];
        res[(r4 - RB * ob + j) * SW + c] = (r4 + j < IH && gv - mean > -13) ? 1 : 0;
      }
    }
    __syncthreads();
    if (ob >= 0 && ob < n_blocks) sink.store(res, img, RB * ob, x0, IH, IW);
    // res is rewritten only after the next iteration's first barrier; the ring rows block b + 1 will
    // overwrite (two blocks ahead of the oldest row any pending window needs) are dead by then
  }
}

// Gaussian taps exactly as cv2.getGaussianKernel(51, 0, CV_32F): double exp, double
// normalisation, cast to float (sigma = 0.3*((51-1)*0.5-1)+0.8 = 8).
static bool g_gauss_ready = false;
static cudaError_t upload_gauss() {
  if (g_gauss_ready) return cudaSuccess;
  double t[KS], sum = 0.0;
  const double sigma = 0.3 * ((KS - 1) * 0.5 - 1) + 0.8;
  for (int i = 0; i < KS; ++i) {
    const double x = i - (KS - 1) * 0.5;
    t[i] = exp(-0.5 / (sigma * sigma) * x * x);
    sum += t[i];
  }
  float h[KS];
  for (int i = 0; i < KS; ++i) h[i] = (float)(t[i] / sum);
  cudaError_t e = cudaMemcpyToSymbol(c_gauss, h, sizeof(h));
  if (e == cudaSuccess) g_gauss_ready = true;
  return e;
}

template <class Src, class Sink>
int launch(bdrc_ctx* ctx, const Src& src, const Sink& sink, int n_img, int IH, int IW,
           cudaStream_t s, bool pdl = false) {
  BDRC_CUDA(ctx, upload_gauss());
  auto kern = adaptive_threshold_kernel<Src, Sink>;
  static bool attr_set = false;   // one flag per <Src, Sink> instantiation
  if (!attr_set) {
    BDRC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    attr_set = true;
  }
  const int rx = ceil_div(IW, SW);
  const long long blocks = (long long)rx * n_img;
  if (blocks == 0 || IH <= 0) return 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = s;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  BDRC_CUDA(ctx, cudaLaunchKernelEx(&cfg, kern, src, sink, IH, IW, rx));
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

}  // namespace adaptive
