// cv2.adaptiveThreshold(ADAPTIVE_THRESH_GAUSSIAN_C, THRESH_BINARY, block=51, C=13) as one fused
// kernel (BDRC/utils.py:796-815): gray source -> fp32 separable 51-tap Gaussian with replicate
// border at the IMAGE (= tile) edge -> mean rounded half-to-even -> gray - mean > -13.
//
// Arithmetic is pinned to oracle/hostops.py:adaptive_threshold: row pass k = 0..50 sequential
// FMA, column pass centre tap then symmetric pairs (x[-k] + x[k]) * w[k] by FMA.
//
// Work decomposition: one CTA produces a 64x64 output region of one image.  The gray halo
// (114x114, clamped to the image) is staged in shared memory as fp32, the row pass writes a
// 114x64 fp32 buffer, the column pass reads it.  Both passes keep a sliding register window so
// shared-memory traffic is ~1 load per 4 FMAs.  The kernel is fp32-FMA bound, not HBM bound:
// ~142 FMA per output pixel against 4 bytes of DRAM traffic.
#pragma once
#include <cmath>

#include "common.cuh"

namespace adaptive {

constexpr int KS = 51, KR = 25;
constexpr int RT = 64;                 // output region edge
constexpr int HALO = RT + 2 * KR;      // 114
constexpr int GSTRIDE = 116;           // padded row stride of the gray tile (float4-aligned)
constexpr int THREADS = 256;
constexpr size_t SMEM_BYTES = (size_t)(HALO * GSTRIDE + HALO * RT) * sizeof(float) + RT * RT;

__constant__ float c_gauss[KS];

// Source policies ------------------------------------------------------------------------------
// RGB page viewed as row-major 512x512 tiles, padded with 255 beyond (Hc, Wc) (utils.py:775-776)
struct SrcPageTiles {
  const uint8_t* page;
  int Hc, Wc, ntx, patch;
  int ih() const { return patch; }
  int iw() const { return patch; }
  __device__ __forceinline__ float gray(int img, int y, int x) const {
    const int ty = img / ntx, tx = img - ty * ntx;
    const int py = ty * patch + y, px = tx * patch + x;
    if (py >= Hc || px >= Wc) return 255.f;
    const uint8_t* p = page + ((size_t)py * Wc + px) * 3;
    const int g = (p[0] * 9798 + p[1] * 19235 + p[2] * 3735 + 16384) >> 15;
    return (float)g;
  }
};

// batch of single-channel u8 images [n][H][W]
struct SrcGrayBatch {
  const uint8_t* img;
  int H, W;
  int ih() const { return H; }
  int iw() const { return W; }
  __device__ __forceinline__ float gray(int i, int y, int x) const {
    return (float)img[((size_t)i * H + y) * W + x];
  }
};

// Sink policies --------------------------------------------------------------------------------
// u8 {0,1} tiles [n][H][W]
struct SinkTiles01 {
  uint8_t* out;
  int H, W;
  __device__ __forceinline__ void store(const uint8_t* res, int img, int y0, int x0, int IH,
                                        int IW) const {
    // res: RT x RT bytes (1 = white). 4 consecutive x per thread -> 32-bit stores when aligned.
    for (int i = threadIdx.x; i < RT * RT / 4; i += THREADS) {
      const int r = i / (RT / 4), c4 = (i % (RT / 4)) * 4;
      const int y = y0 + r, x = x0 + c4;
      if (y >= IH || x >= IW) continue;
      uint8_t* o = out + ((size_t)img * H + y) * W + x;
      if (x + 3 < IW && (W & 3) == 0) {
        *reinterpret_cast<uint32_t*>(o) = *reinterpret_cast<const uint32_t*>(res + r * RT + c4);
      } else {
        for (int j = 0; j < 4 && x + j < IW; ++j) o[j] = res[r * RT + c4 + j];
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
  __device__ __forceinline__ void store(const uint8_t* res, int img, int y0, int x0, int IH,
                                        int IW) const {
    if (lines) {
      // lanes run along y (the channel axis of the output) so stores are contiguous
      for (int i = threadIdx.x; i < RT * RT; i += THREADS) {
        const int c = i / RT, r = i % RT;   // c = x offset, r = y offset
        const int y = y0 + r, x = x0 + c;
        if (y >= IH || x >= IW) continue;
        lines[((size_t)img * W + x) * Cpad + y] = __float2bfloat16(res[r * RT + c] ? 1.f : -1.f);
      }
    }
    if (lines_f32) {
      for (int i = threadIdx.x; i < RT * RT; i += THREADS) {
        const int r = i / RT, c = i % RT;
        const int y = y0 + r, x = x0 + c;
        if (y >= IH || x >= IW) continue;
        lines_f32[((size_t)img * H + y) * W + x] = res[r * RT + c] ? 1.f : -1.f;
      }
    }
  }
};

template <class Src, class Sink>
__global__ void __launch_bounds__(THREADS)
adaptive_threshold_kernel(const Src src, const Sink sink, int IH, int IW, int regions_x,
                          int regions_y) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  float* g = reinterpret_cast<float*>(smem_raw);       // [HALO][GSTRIDE]
  float* rowb = g + HALO * GSTRIDE;                     // [HALO][RT]
  uint8_t* res = reinterpret_cast<uint8_t*>(rowb + HALO * RT);  // [RT][RT]

  int b = blockIdx.x;
  const int rx = b % regions_x; b /= regions_x;
  const int ry = b % regions_y;
  const int img = b / regions_y;
  const int x0 = rx * RT, y0 = ry * RT;

  // ---- stage gray halo (replicate border at the image edge) ----
  for (int i = threadIdx.x; i < HALO * HALO; i += THREADS) {
    const int r = i / HALO, c = i - r * HALO;
    const int y = min(max(y0 + r - KR, 0), IH - 1);
    const int x = min(max(x0 + c - KR, 0), IW - 1);
    g[r * GSTRIDE + c] = src.gray(img, y, x);
  }
  __syncthreads();

  // ---- row pass: rowb[r][c] = sum_k g[r][c + k] * w[k], k ascending, FMA ----
  for (int it = threadIdx.x; it < HALO * (RT / 4); it += THREADS) {
    const int r = it / (RT / 4), c = (it - r * (RT / 4)) * 4;
    const float* gp = g + r * GSTRIDE + c;
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
    *reinterpret_cast<float4*>(rowb + r * RT + c) = make_float4(a0, a1, a2, a3);
  }
  __syncthreads();

  // ---- column pass + threshold: 4 consecutive rows per thread, lanes along x ----
  for (int it = threadIdx.x; it < (RT / 4) * RT; it += THREADS) {
    const int c = it % RT, r4 = (it / RT) * 4;   // output rows r4 .. r4+3 (region coords)
    // output row r uses rowb rows (r + KR - k) and (r + KR + k): 54-row register window
    const float* rp = rowb + r4 * RT + c;
    float v[2 * KR + 4];
#pragma unroll
    for (int i = 0; i < 2 * KR + 4; ++i) v[i] = rp[i * RT];
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
      const int gv = (int)g[(r4 + j + KR) * GSTRIDE + c + KR];
      res[(r4 + j) * RT + c] = (gv - mean > -13) ? 1 : 0;
    }
  }
  __syncthreads();
  sink.store(res, img, y0, x0, IH, IW);
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
           cudaStream_t s) {
  BDRC_CUDA(ctx, upload_gauss());
  auto kern = adaptive_threshold_kernel<Src, Sink>;
  BDRC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SMEM_BYTES));
  const int rx = ceil_div(IW, RT), ry = ceil_div(IH, RT);
  const long long blocks = (long long)rx * ry * n_img;
  if (blocks == 0) return 0;
  kern<<<(unsigned)blocks, THREADS, SMEM_BYTES, s>>>(src, sink, IH, IW, rx, ry);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

}  // namespace adaptive
