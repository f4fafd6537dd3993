// Line-level image stages of the hot path:
//   a-5  extract_line_images / get_line_image / extract_line / mask_n_crop
//        (BDRC/line_detection.py:79-104, 267-359): rect dilation of the filled-contour mask,
//        AND with the page, deletion of all-zero rows / columns, ragged crop gather
//   a-4  rotate_from_angle = cv2.warpAffine, fixed-point bilinear (BDRC/line_detection.py:62-76)
//   a-6  bilinear remap with explicit or thin-plate-spline coordinates
//        (BDRC/image_dewarping.py:55-72, order=1 per north_star)
//
// The per-line contour is rasterised on the host with cv2.drawContours into its bounding-box ROI
// (tens of kB) exactly as the reference does on a full-page canvas; everything that touches
// page-sized data runs here.  Dilation by a kw x kh rectangle is evaluated in O(1) per pixel from
// a summed-area table of the ROI mask.
#include "common.cuh"

namespace {

struct LineGeo {
  int rx, ry, rw, rh;          // ROI of the filled-contour mask (page coords)
  int kw, kh;                  // structuring element (cv2.getStructuringElement(MORPH_RECT, (kw, kh)))
  int bx0, by0, bw, bh;        // dilated bounding box clipped to the page
  long long sat_off;           // int32 offset of this line's (rh+1) x (rw+1) summed-area table
  long long mask_off;          // byte offset of the ROI mask
  long long row_off, col_off;  // offsets into the per-line row / column scratch arrays
};

__host__ __device__ inline void dilated_bbox(LineGeo& g, int H, int W) {
  const int ax = g.kw / 2, ay = g.kh / 2;
  int x0 = g.rx - (g.kw - 1 - ax), x1 = g.rx + g.rw - 1 + ax;
  int y0 = g.ry - (g.kh - 1 - ay), y1 = g.ry + g.rh - 1 + ay;
  x0 = x0 < 0 ? 0 : x0; y0 = y0 < 0 ? 0 : y0;
  x1 = x1 > W - 1 ? W - 1 : x1; y1 = y1 > H - 1 ? H - 1 : y1;
  g.bx0 = x0; g.by0 = y0;
  g.bw = x1 >= x0 ? x1 - x0 + 1 : 0;
  g.bh = y1 >= y0 ? y1 - y0 + 1 : 0;
}

// cv2.dilate(mask, rect(kw,kh)) at page pixel (y, x): any set ROI pixel inside the window
// rows [y - ay, y - ay + kh - 1], cols [x - ax, x - ax + kw - 1] (anchor = centre, border = 0)
__device__ __forceinline__ bool dilated_at(const LineGeo& g, const int* __restrict__ sat, int y, int x) {
  const int ax = g.kw / 2, ay = g.kh / 2;
  int y0 = y - ay - g.ry, y1 = y0 + g.kh - 1;
  int x0 = x - ax - g.rx, x1 = x0 + g.kw - 1;
  y0 = max(y0, 0); x0 = max(x0, 0);
  y1 = min(y1, g.rh - 1); x1 = min(x1, g.rw - 1);
  if (y0 > y1 || x0 > x1) return false;
  const int* s = sat + g.sat_off;
  const int p = g.rw + 1;
  return s[(y1 + 1) * p + x1 + 1] - s[y0 * p + x1 + 1] - s[(y1 + 1) * p + x0] + s[y0 * p + x0] > 0;
}

// ---- summed-area table of each ROI mask: see sat_rows_strided / sat_cols_strided below ------

// ---- row statistics: keep row iff every channel has a non-zero masked pixel -------------------
// one CTA per (line-or-candidate, dilated row)
__global__ void __launch_bounds__(256)
row_keep_kernel(const uint8_t* __restrict__ page, int H, int W, int C, const LineGeo* __restrict__ geo,
                const int* __restrict__ sat, uint8_t* __restrict__ rowkeep) {
  const LineGeo g = geo[blockIdx.y];
  const int dy = blockIdx.x;
  if (dy >= g.bh) return;
  const int y = g.by0 + dy;
  unsigned flags = 0;
  for (int dx = threadIdx.x; dx < g.bw; dx += blockDim.x) {
    const int x = g.bx0 + dx;
    if (!dilated_at(g, sat, y, x)) continue;
    const uint8_t* px = page + ((size_t)y * W + x) * C;
    for (int c = 0; c < C; ++c)
      if (px[c]) flags |= 1u << c;
  }
  __shared__ unsigned sh;
  if (threadIdx.x == 0) sh = 0;
  __syncthreads();
  if (flags) atomicOr(&sh, flags);
  __syncthreads();
  if (threadIdx.x == 0) rowkeep[g.row_off + dy] = (sh == (1u << C) - 1) ? 1 : 0;
}

// per line: compact kept rows (rowmap[i] = dy of the i-th kept row) and count them
__global__ void __launch_bounds__(256)
compact_kernel(const uint8_t* __restrict__ keep, const LineGeo* __restrict__ geo, int is_col,
               int* __restrict__ map, int* __restrict__ counts) {
  const LineGeo g = geo[blockIdx.x];
  const long long off = is_col ? g.col_off : g.row_off;
  const int n = is_col ? g.bw : g.bh;
  __shared__ int warp_tot[8];
  __shared__ int base;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  for (int i0 = 0; i0 < n; i0 += 256) {
    const int i = i0 + threadIdx.x;
    const int k = (i < n && keep[off + i]) ? 1 : 0;
    int v = k;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    if (lane == 31) warp_tot[warp] = v;
    __syncthreads();
    int pre = 0;
    for (int w = 0; w < warp; ++w) pre += warp_tot[w];
    if (k) map[off + base + pre + v - 1] = i;
    __syncthreads();
    if (threadIdx.x == 255) base += pre + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) counts[blockIdx.x] = base;
}

// crop heights of many (line, candidate SE) pairs: counts[] as produced by compact_kernel on rows
// column statistics over the kept rows only (np.delete of rows happens first: line_detection.py:97-102)
__global__ void col_keep_kernel(const uint8_t* __restrict__ page, int H, int W, int C,
                                const LineGeo* __restrict__ geo, const int* __restrict__ sat,
                                const int* __restrict__ rowmap, const int* __restrict__ nrows,
                                uint8_t* __restrict__ colkeep) {
  const LineGeo g = geo[blockIdx.y];
  const int dx = blockIdx.x * blockDim.x + threadIdx.x;
  if (dx >= g.bw) return;
  const int x = g.bx0 + dx;
  const int hr = nrows[blockIdx.y];
  unsigned flags = 0;
  const unsigned full = (1u << C) - 1;
  for (int i = 0; i < hr && flags != full; ++i) {
    const int y = g.by0 + rowmap[g.row_off + i];
    if (!dilated_at(g, sat, y, x)) continue;
    const uint8_t* px = page + ((size_t)y * W + x) * C;
    for (int c = 0; c < C; ++c)
      if (px[c]) flags |= 1u << c;
  }
  colkeep[g.col_off + dx] = flags == full ? 1 : 0;
}

__global__ void crop_layout_kernel(const int* __restrict__ nrows, const int* __restrict__ ncols, int L, int C,
                                   int32_t* __restrict__ crop_hw, int64_t* __restrict__ crop_off,
                                   long long capacity, int* __restrict__ overflow) {
  if (blockIdx.x || threadIdx.x) return;
  long long off = 0;
  for (int l = 0; l < L; ++l) {
    const int h = nrows[l], w = h > 0 ? ncols[l] : 0;
    crop_hw[2 * l] = h;
    crop_hw[2 * l + 1] = w;
    crop_off[l] = off;
    off += ((long long)h * w * C + 15) / 16 * 16;
  }
  *overflow = off > capacity ? 1 : 0;
}

__global__ void gather_kernel(const uint8_t* __restrict__ page, int H, int W, int C,
                              const LineGeo* __restrict__ geo, const int* __restrict__ sat,
                              const int* __restrict__ rowmap, const int* __restrict__ colmap,
                              const int32_t* __restrict__ crop_hw, const int64_t* __restrict__ crop_off,
                              const int* __restrict__ overflow, uint8_t* __restrict__ crops) {
  if (*overflow) return;
  const int l = blockIdx.z;
  const LineGeo g = geo[l];
  const int h = crop_hw[2 * l], w = crop_hw[2 * l + 1];
  const int r = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= h || c >= w) return;
  const int y = g.by0 + rowmap[g.row_off + r], x = g.bx0 + colmap[g.col_off + c];
  uint8_t* o = crops + crop_off[l] + ((size_t)r * w + c) * C;
  if (dilated_at(g, sat, y, x)) {
    const uint8_t* px = page + ((size_t)y * W + x) * C;
    for (int k = 0; k < C; ++k) o[k] = px[k];
  } else {
    for (int k = 0; k < C; ++k) o[k] = 0;
  }
}

// ---- cv2.warpAffine (INTER_LINEAR, BORDER_CONSTANT 0) -----------------------------------------
struct AffineInv { double A11, A12, A21, A22, b1, b2; };

__global__ void warp_affine_kernel(const uint8_t* __restrict__ src, int H, int W, int C, AffineInv m,
                                   uint8_t* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  // OpenCV: AB_BITS = 10, INTER_BITS = 5, round_delta = 16, 15-bit bilinear weights
  const int adelta = __double2int_rn(m.A11 * (double)x * 1024.0);
  const int bdelta = __double2int_rn(m.A21 * (double)x * 1024.0);
  const int X0 = __double2int_rn((m.A12 * (double)y + m.b1) * 1024.0) + 16;
  const int Y0 = __double2int_rn((m.A22 * (double)y + m.b2) * 1024.0) + 16;
  const int X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  const int sx = X >> 5, sy = Y >> 5, fx = X & 31, fy = Y & 31;
  const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32;
  const int w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
  const bool y0ok = sy >= 0 && sy < H, y1ok = sy + 1 >= 0 && sy + 1 < H;
  const bool x0ok = sx >= 0 && sx < W, x1ok = sx + 1 >= 0 && sx + 1 < W;
  uint8_t* o = dst + ((size_t)y * W + x) * C;
  for (int c = 0; c < C; ++c) {
    const int v00 = (y0ok && x0ok) ? src[((size_t)sy * W + sx) * C + c] : 0;
    const int v01 = (y0ok && x1ok) ? src[((size_t)sy * W + sx + 1) * C + c] : 0;
    const int v10 = (y1ok && x0ok) ? src[((size_t)(sy + 1) * W + sx) * C + c] : 0;
    const int v11 = (y1ok && x1ok) ? src[((size_t)(sy + 1) * W + sx + 1) * C + c] : 0;
    const int v = (v00 * w00 + v01 * w01 + v10 * w10 + v11 * w11 + (1 << 14)) >> 15;
    o[c] = (uint8_t)min(max(v, 0), 255);
  }
}

// ---- scipy.ndimage.map_coordinates(order=1, mode='constant', cval=0) --------------------------
__device__ __forceinline__ void sample_bilinear_const(const uint8_t* __restrict__ src, int H, int W, int C,
                                                      double r, double c, uint8_t* o) {
  if (!(r >= 0.0 && r <= (double)(H - 1) && c >= 0.0 && c <= (double)(W - 1))) {
    for (int k = 0; k < C; ++k) o[k] = 0;
    return;
  }
  const int r0 = (int)floor(r), c0 = (int)floor(c);
  const double fr = r - r0, fc = c - c0;
  const int r1 = min(r0 + 1, H - 1), c1 = min(c0 + 1, W - 1);
  for (int k = 0; k < C; ++k) {
    const double v = (1.0 - fr) * ((1.0 - fc) * src[((size_t)r0 * W + c0) * C + k] + fc * src[((size_t)r0 * W + c1) * C + k]) +
                     fr * ((1.0 - fc) * src[((size_t)r1 * W + c0) * C + k] + fc * src[((size_t)r1 * W + c1) * C + k]);
    o[k] = (uint8_t)min(max((int)floor(v + 0.5), 0), 255);
  }
}

__global__ void remap_kernel(const uint8_t* __restrict__ src, int H, int W, int C,
                             const float* __restrict__ coords, uint8_t* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  const size_t i = (size_t)y * W + x;
  sample_bilinear_const(src, H, W, C, (double)coords[i], (double)coords[(size_t)H * W + i], dst + i * C);
}


// ---- scipy.ndimage.map_coordinates(order=3, mode='constant', cval=0): the sampling the reference's run_tps
// intends (image_dewarping.py:66-72 calls map_coordinates with its defaults) ---------------------------------
// Two steps, as in scipy: (1) spline_filter: the image becomes cubic B-spline coefficients by a causal +
// anticausal first-order recursion with pole z = sqrt(3) - 2 along each axis (axis 0 first), mirror boundary;
// (2) every output sample is a 4 x 4 weighted sum of coefficients, taps mirrored at the border, 0 outside
// [0, n-1].  All arithmetic is fp64 and unfused (scipy's C is compiled without FMA contraction), so the
// interpolation step is bit-identical to scipy given the same coefficients; the coefficients agree to a few ulp.
constexpr double SPL_Z = -0.26794919243112270647;   // sqrt(3) - 2
constexpr int SPL_INIT_TERMS = 80;                  // |z|^80 < 2e-46: the causal start-up sum has converged

// A line is cut into segments of SPL_SEG samples, one thread per (segment, line).  The recursions forget their
// past at |z|^k: a thread starts SPL_WARM samples before its segment (z^64 = 3e-37, far below one fp64 ulp) and
// runs the causal sweep SPL_WARM samples past its end, keeping that tail in shared memory as the run-in of the
// anticausal sweep; segments at the ends of the line use scipy's exact mirror start-up formulas.  This gives
// n / SPL_SEG times the parallelism of one thread per line (74 k threads per axis on a 2048 x 3072 x 3 page).
constexpr int SPL_SEG = 256, SPL_WARM = 64, SPL_THREADS = 64;

// element i of line l lives at base[(l / inner) * outer + (l % inner) + i * stride]
template <typename Tin>
__global__ void __launch_bounds__(SPL_THREADS)
spline3_filter_kernel(const Tin* in, double* out, int n_lines, int n, int inner, size_t outer, size_t stride) {
  __shared__ double tail_s[SPL_WARM * SPL_THREADS];
  const long long t = (long long)blockIdx.x * SPL_THREADS + threadIdx.x;
  const int n_seg = (n + SPL_SEG - 1) / SPL_SEG;
  if (t >= (long long)n_lines * n_seg) return;
  const int l = (int)(t % n_lines), seg = (int)(t / n_lines);
  const size_t base = (size_t)(l / inner) * outer + (size_t)(l % inner);
  in += base;
  out += base;
  double* tail = tail_s + threadIdx.x;            // tail[k * SPL_THREADS]: conflict-free across the block
  const double z = SPL_Z;
  const double gain = __dmul_rn(1.0 - z, 1.0 - 1.0 / z);
  if (n == 1) {
    out[0] = (double)in[0];
    return;
  }
  const int s = seg * SPL_SEG, e = min(s + SPL_SEG, n);
  // ---- causal sweep: c+[i] = gain * x[i] + z * c+[i-1] ----
  double prev, pprev = 0.0;
  int i;
  if (s - SPL_WARM <= 0) {
    // exact start (ni_splines.c:_init_causal_mirror), the sum truncated where z^i has vanished
    const double z_n_1 = pow(z, (double)(n - 1));
    double c0 = __dadd_rn(__dmul_rn((double)in[0], gain), __dmul_rn(z_n_1, __dmul_rn((double)in[(size_t)(n - 1) * stride], gain)));
    double z_i = z;
    const int terms = min(n - 2, SPL_INIT_TERMS);
    for (int k = 1; k <= terms; ++k) {
      const double a = __dmul_rn((double)in[(size_t)k * stride], gain);
      const double b = __dmul_rn((double)in[(size_t)(n - 1 - k) * stride], gain);
      c0 = __dadd_rn(c0, __dmul_rn(z_i, __dadd_rn(a, __dmul_rn(z_n_1, b))));
      z_i = __dmul_rn(z_i, z);
    }
    prev = c0 / __dadd_rn(1.0, -__dmul_rn(z_n_1, z_n_1));
    if (s == 0) out[0] = prev;
    i = 1;
  } else {
    prev = __dmul_rn((double)in[(size_t)(s - SPL_WARM) * stride], gain);   // history dropped: error z^SPL_WARM
    i = s - SPL_WARM + 1;
  }
  for (; i < s; ++i) {
    pprev = prev;
    prev = __dadd_rn(__dmul_rn((double)in[(size_t)i * stride], gain), __dmul_rn(z, prev));
  }
#pragma unroll 4
  for (; i < e; ++i) {
    pprev = prev;
    prev = __dadd_rn(__dmul_rn((double)in[(size_t)i * stride], gain), __dmul_rn(z, prev));
    out[(size_t)i * stride] = prev;
  }
  const int te = min(e + SPL_WARM, n);
  for (; i < te; ++i) {
    pprev = prev;
    prev = __dadd_rn(__dmul_rn((double)in[(size_t)i * stride], gain), __dmul_rn(z, prev));
    tail[(i - e) * SPL_THREADS] = prev;
  }
  // ---- anticausal sweep: c[i] = z * (c[i+1] - c+[i]) ----
  double nxt;
  int j;
  if (te == n) {
    // exact end (_init_anticausal_mirror): prev = c+[n-1], pprev = c+[n-2]
    nxt = __dmul_rn(__dadd_rn(__dmul_rn(z, pprev), prev), z) / __dadd_rn(__dmul_rn(z, z), -1.0);
    if (n - 1 < e) out[(size_t)(n - 1) * stride] = nxt;
    j = n - 2;
  } else {
    nxt = 0.0;                                                              // run-in: error z^SPL_WARM
    j = te - 1;
  }
  for (; j >= e; --j) nxt = __dmul_rn(z, __dadd_rn(nxt, -tail[(j - e) * SPL_THREADS]));
#pragma unroll 4
  for (; j >= s; --j) {
    nxt = __dmul_rn(z, __dadd_rn(nxt, -out[(size_t)j * stride]));
    out[(size_t)j * stride] = nxt;
  }
}

__device__ __forceinline__ void spline3_weights(double x, double* w) {
  // ni_splines.c:get_spline_interpolation_weights, order 3
  x = __dadd_rn(x, -floor(x));
  const double y = x, z = __dadd_rn(1.0, -x);
  w[1] = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(y, y), __dadd_rn(y, -2.0)), 3.0), 4.0) / 6.0;
  w[2] = __dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(z, z), __dadd_rn(z, -2.0)), 3.0), 4.0) / 6.0;
  w[0] = __dmul_rn(__dmul_rn(z, z), z) / 6.0;
  w[3] = __dadd_rn(__dadd_rn(__dadd_rn(1.0, -w[0]), -w[1]), -w[2]);
}

__device__ __forceinline__ int spline_mirror(int idx, int n) {
  // ni_interpolation.c border mapping of the filter taps for mode 'constant' (whole-sample mirror)
  if (n <= 1) return 0;
  const int s2 = 2 * n - 2;
  if (idx < 0) {
    idx = s2 * (-idx / s2) + idx;
    return idx <= 1 - n ? idx + s2 : -idx;
  }
  if (idx >= n) {
    idx -= s2 * (idx / s2);
    if (idx >= n) idx = s2 - idx;
  }
  return idx;
}

__device__ __forceinline__ void sample_cubic_const(const double* __restrict__ coef, int H, int W, int C, double r,
                                                   double c, uint8_t* o) {
  if (!(r >= 0.0 && r <= (double)(H - 1) && c >= 0.0 && c <= (double)(W - 1))) {
    for (int k = 0; k < C; ++k) o[k] = 0;
    return;
  }
  const int sr = (int)floor(r) - 1, sc = (int)floor(c) - 1;
  double wr[4], wc[4];
  spline3_weights(r, wr);
  spline3_weights(c, wc);
  int ri[4], ci[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    ri[i] = spline_mirror(sr + i, H);
    ci[i] = spline_mirror(sc + i, W);
  }
  for (int k = 0; k < C; ++k) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        t = __dadd_rn(t, __dmul_rn(__dmul_rn(coef[((size_t)ri[i] * W + ci[j]) * C + k], wr[i]), wc[j]));
    // ni_interpolation.c CASE_INTERP_OUT_UINT: t > 0 ? t + 0.5 : 0, clamped, truncated
    t = t > 0.0 ? __dadd_rn(t, 0.5) : 0.0;
    o[k] = (uint8_t)(t > 255.0 ? 255.0 : t);
  }
}

__global__ void remap_cubic_kernel(const double* __restrict__ coef, int H, int W, int C,
                                   const double* __restrict__ coords, uint8_t* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  const size_t i = (size_t)y * W + x;
  sample_cubic_const(coef, H, W, C, coords[i], coords[(size_t)H * W + i], dst + i * C);
}

constexpr int TPS_MAX_CTRL = 64;
struct TpsParams {
  int n;
  double ctrl[TPS_MAX_CTRL][2];
  double par[TPS_MAX_CTRL + 3][2];
};

// thin-plate-spline transform of the output grid index (row, col), then bilinear (CUBIC = false) or cubic
// B-spline (CUBIC = true, `src` then holds fp64 spline coefficients) sampling
// (tps.ThinPlateSpline.transform: [U(|p - c_j|), 1, p] @ parameters, U(r) = r^2 log r)
template <bool CUBIC>
__global__ void tps_remap_kernel(const void* __restrict__ src, int H, int W, int C,
                                 const TpsParams* __restrict__ tp, uint8_t* __restrict__ dst) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W) return;
  const int n = tp->n;
  const double pr = (double)y, pc = (double)x;
  double orow = tp->par[n][0] + pr * tp->par[n + 1][0] + pc * tp->par[n + 2][0];
  double ocol = tp->par[n][1] + pr * tp->par[n + 1][1] + pc * tp->par[n + 2][1];
  for (int j = 0; j < n; ++j) {
    const double dr = pr - tp->ctrl[j][0], dc = pc - tp->ctrl[j][1];
    const double d2 = dr * dr + dc * dc;
    const double u = d2 > 0.0 ? 0.5 * d2 * log(d2) : 0.0;   // r^2 log r = 0.5 r^2 log r^2
    orow += u * tp->par[j][0];
    ocol += u * tp->par[j][1];
  }
  if (CUBIC)
    sample_cubic_const(reinterpret_cast<const double*>(src), H, W, C, orow, ocol, dst + ((size_t)y * W + x) * C);
  else
    sample_bilinear_const(reinterpret_cast<const uint8_t*>(src), H, W, C, orow, ocol, dst + ((size_t)y * W + x) * C);
}

// workspace carve-up shared by the two extract entry points
struct ExtractWs {
  LineGeo* geo; int* sat; uint8_t* rowkeep; uint8_t* colkeep; int* rowmap; int* colmap;
  int* nrows; int* ncols; int* overflow;
};

size_t extract_ws_bytes(int H, int W, int n_items, long long sat_elems) {
  size_t b = 0;
  b += align_up((size_t)n_items * sizeof(LineGeo), 256);
  b += align_up((size_t)sat_elems * sizeof(int), 256);
  b += align_up((size_t)n_items * H, 256) + align_up((size_t)n_items * W, 256);
  b += align_up((size_t)n_items * H * sizeof(int), 256) + align_up((size_t)n_items * W * sizeof(int), 256);
  b += 3 * align_up((size_t)n_items * sizeof(int), 256) + 256;
  return b;
}

bool carve(ExtractWs* w, void* ws, size_t bytes, int H, int W, int n_items, long long sat_elems) {
  if (bytes < extract_ws_bytes(H, W, n_items, sat_elems)) return false;
  uint8_t* p = reinterpret_cast<uint8_t*>(ws);
  auto take = [&](size_t n) { uint8_t* r = p; p += align_up(n, 256); return r; };
  w->geo = reinterpret_cast<LineGeo*>(take((size_t)n_items * sizeof(LineGeo)));
  w->sat = reinterpret_cast<int*>(take((size_t)sat_elems * sizeof(int)));
  w->rowkeep = take((size_t)n_items * H);
  w->colkeep = take((size_t)n_items * W);
  w->rowmap = reinterpret_cast<int*>(take((size_t)n_items * H * sizeof(int)));
  w->colmap = reinterpret_cast<int*>(take((size_t)n_items * W * sizeof(int)));
  w->nrows = reinterpret_cast<int*>(take((size_t)n_items * sizeof(int)));
  w->ncols = reinterpret_cast<int*>(take((size_t)n_items * sizeof(int)));
  w->overflow = reinterpret_cast<int*>(take(sizeof(int)));
  return true;
}

// builds the per-item geometry (one item = one (line, SE) pair), uploads it and fills the SATs
int prepare_items(bdrc_ctx* ctx, int H, int W, int L, const int32_t* roi_host, const int64_t* mask_off_host,
                  int n_cand, const int32_t* kw_host, const int32_t* kh_host, const uint8_t* masks,
                  ExtractWs& w, std::vector<LineGeo>& items, int* max_bw, int* max_bh, int* max_rw,
                  int* max_rh, cudaStream_t s) {
  // SATs are per LINE (shared by that line's candidates)
  std::vector<long long> sat_off(L);
  long long so = 0;
  for (int l = 0; l < L; ++l) {
    sat_off[l] = so;
    so += (long long)(roi_host[4 * l + 2] + 1) * (roi_host[4 * l + 3] + 1);
  }
  items.resize((size_t)L * n_cand);
  *max_bw = *max_bh = *max_rw = *max_rh = 0;
  for (int l = 0; l < L; ++l)
    for (int j = 0; j < n_cand; ++j) {
      LineGeo g;
      g.rx = roi_host[4 * l]; g.ry = roi_host[4 * l + 1]; g.rw = roi_host[4 * l + 2]; g.rh = roi_host[4 * l + 3];
      g.kw = kw_host[l * n_cand + j]; g.kh = kh_host[l * n_cand + j];
      if (g.kw < 1 || g.kh < 1 || g.rw < 1 || g.rh < 1)
        return bdrc_fail(ctx, BDRC_ERR_ARG, "extract_lines: empty ROI or structuring element (line %d)", l);
      dilated_bbox(g, H, W);
      g.sat_off = sat_off[l];
      g.mask_off = mask_off_host[l];
      const size_t it = (size_t)l * n_cand + j;
      g.row_off = (long long)it * H;
      g.col_off = (long long)it * W;
      items[it] = g;
      *max_bw = g.bw > *max_bw ? g.bw : *max_bw;
      *max_bh = g.bh > *max_bh ? g.bh : *max_bh;
      *max_rw = g.rw > *max_rw ? g.rw : *max_rw;
      *max_rh = g.rh > *max_rh ? g.rh : *max_rh;
    }
  BDRC_CUDA(ctx, cudaMemcpyAsync(w.geo, items.data(), items.size() * sizeof(LineGeo), cudaMemcpyHostToDevice, s));
  // SAT build uses the first candidate of every line (stride n_cand in the item array): launch on a
  // compacted view by passing a temporary array of per-line items at the tail of geo is avoided --
  // candidates of a line share rx/ry/rw/rh/sat_off/mask_off, so indexing item l*n_cand is enough.
  return 0;
}

__global__ void sat_rows_strided(const uint8_t* masks, const LineGeo* geo, int n_cand, int* sat) {
  // thin wrapper so that blockIdx.y indexes lines, not items
  const LineGeo g = geo[(size_t)blockIdx.y * n_cand];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int p = g.rw + 1;
  int* s = sat + g.sat_off;
  for (int r = blockIdx.x * nw + warp; r <= g.rh; r += gridDim.x * nw) {
    if (r == 0) {
      for (int c = lane; c <= g.rw; c += 32) s[c] = 0;
      continue;
    }
    const uint8_t* m = masks + g.mask_off + (size_t)(r - 1) * g.rw;
    int carry = 0;
    if (lane == 0) s[r * p] = 0;
    for (int c0 = 0; c0 < g.rw; c0 += 32) {
      const int c = c0 + lane;
      int v = (c < g.rw && m[c] != 0) ? 1 : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
      }
      if (c < g.rw) s[r * p + c + 1] = carry + v;
      carry += __shfl_sync(0xffffffffu, v, 31);
    }
  }
}

__global__ void sat_cols_strided(const LineGeo* geo, int n_cand, int* sat) {
  const LineGeo g = geo[(size_t)blockIdx.y * n_cand];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > g.rw) return;
  int* s = sat + g.sat_off;
  const int p = g.rw + 1;
  int acc = 0;
  for (int r = 1; r <= g.rh; ++r) {
    acc += s[r * p + c];
    s[r * p + c] = acc;
  }
}

long long sat_elems_of(const int32_t* roi_host, int L) {
  long long so = 0;
  for (int l = 0; l < L; ++l) so += (long long)(roi_host[4 * l + 2] + 1) * (roi_host[4 * l + 3] + 1);
  return so;
}

}  // namespace

extern "C" {

size_t bdrc_extract_lines_workspace_bytes(int H, int W, int n_items, int64_t sat_elems) {
  return extract_ws_bytes(H, W, n_items, sat_elems);
}

int bdrc_line_crop_heights(bdrc_ctx* ctx, const uint8_t* page, int H, int W, int C, const uint8_t* masks,
                           const int32_t* roi_host, const int64_t* mask_off_host, int L, int n_cand,
                           const int32_t* kw_host, const int32_t* kh_host, int32_t* heights,
                           void* workspace, size_t workspace_bytes, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  if (L == 0 || n_cand == 0) return 0;
  BDRC_REQUIRE(ctx, page && masks && roi_host && mask_off_host && kw_host && kh_host && heights && (C == 1 || C == 3),
               "line_crop_heights: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const long long se = sat_elems_of(roi_host, L);
  ExtractWs w;
  BDRC_REQUIRE(ctx, carve(&w, workspace, workspace_bytes, H, W, L * n_cand, se), "line_crop_heights: workspace too small");
  std::vector<LineGeo> items;
  int mbw, mbh, mrw, mrh;
  BDRC_TRY(prepare_items(ctx, H, W, L, roi_host, mask_off_host, n_cand, kw_host, kh_host, masks, w, items, &mbw,
                         &mbh, &mrw, &mrh, s));
  sat_rows_strided<<<dim3(ceil_div(mrh + 1, 8), L), 256, 0, s>>>(masks, w.geo, n_cand, w.sat);
  BDRC_CHECK_LAUNCH(ctx);
  sat_cols_strided<<<dim3(ceil_div(mrw + 1, 256), L), 256, 0, s>>>(w.geo, n_cand, w.sat);
  BDRC_CHECK_LAUNCH(ctx);
  if (mbh > 0) {
    row_keep_kernel<<<dim3(mbh, L * n_cand), 256, 0, s>>>(page, H, W, C, w.geo, w.sat, w.rowkeep);
    BDRC_CHECK_LAUNCH(ctx);
  }
  compact_kernel<<<L * n_cand, 256, 0, s>>>(w.rowkeep, w.geo, 0, w.rowmap, heights);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

int bdrc_extract_lines(bdrc_ctx* ctx, const uint8_t* page, int H, int W, int C, const uint8_t* masks,
                       const int32_t* roi_host, const int64_t* mask_off_host, int L, const int32_t* kw_host,
                       const int32_t* kh_host, uint8_t* crops, size_t crops_capacity, int32_t* crop_hw,
                       int64_t* crop_off, void* workspace, size_t workspace_bytes, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  if (L == 0) return 0;
  BDRC_REQUIRE(ctx, page && masks && roi_host && mask_off_host && kw_host && kh_host && crops && crop_hw && crop_off &&
                        (C == 1 || C == 3), "extract_lines: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const long long se = sat_elems_of(roi_host, L);
  ExtractWs w;
  BDRC_REQUIRE(ctx, carve(&w, workspace, workspace_bytes, H, W, L, se), "extract_lines: workspace too small");
  std::vector<LineGeo> items;
  int mbw, mbh, mrw, mrh;
  BDRC_TRY(prepare_items(ctx, H, W, L, roi_host, mask_off_host, 1, kw_host, kh_host, masks, w, items, &mbw, &mbh,
                         &mrw, &mrh, s));
  sat_rows_strided<<<dim3(ceil_div(mrh + 1, 8), L), 256, 0, s>>>(masks, w.geo, 1, w.sat);
  BDRC_CHECK_LAUNCH(ctx);
  sat_cols_strided<<<dim3(ceil_div(mrw + 1, 256), L), 256, 0, s>>>(w.geo, 1, w.sat);
  BDRC_CHECK_LAUNCH(ctx);
  if (mbh > 0 && mbw > 0) {
    row_keep_kernel<<<dim3(mbh, L), 256, 0, s>>>(page, H, W, C, w.geo, w.sat, w.rowkeep);
    BDRC_CHECK_LAUNCH(ctx);
  }
  compact_kernel<<<L, 256, 0, s>>>(w.rowkeep, w.geo, 0, w.rowmap, w.nrows);
  BDRC_CHECK_LAUNCH(ctx);
  if (mbw > 0) {
    col_keep_kernel<<<dim3(ceil_div(mbw, 128), L), 128, 0, s>>>(page, H, W, C, w.geo, w.sat, w.rowmap, w.nrows,
                                                               w.colkeep);
    BDRC_CHECK_LAUNCH(ctx);
  }
  compact_kernel<<<L, 256, 0, s>>>(w.colkeep, w.geo, 1, w.colmap, w.ncols);
  BDRC_CHECK_LAUNCH(ctx);
  crop_layout_kernel<<<1, 1, 0, s>>>(w.nrows, w.ncols, L, C, crop_hw, crop_off, (long long)crops_capacity, w.overflow);
  BDRC_CHECK_LAUNCH(ctx);
  if (mbh > 0 && mbw > 0) {
    gather_kernel<<<dim3(ceil_div(mbw, 128), mbh, L), 128, 0, s>>>(page, H, W, C, w.geo, w.sat, w.rowmap, w.colmap,
                                                                  crop_hw, crop_off, w.overflow, crops);
    BDRC_CHECK_LAUNCH(ctx);
  }
  return 0;
}

int bdrc_warp_affine_u8(bdrc_ctx* ctx, const uint8_t* src, int H, int W, int C, const double* M, uint8_t* dst,
                        void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  BDRC_REQUIRE(ctx, src && dst && M && H > 0 && W > 0 && C >= 1 && C <= 4, "warp_affine_u8: bad arguments");
  // invert the 2x3 matrix exactly like cv::warpAffine (without WARP_INVERSE_MAP)
  double D = M[0] * M[4] - M[1] * M[3];
  D = D != 0.0 ? 1.0 / D : 0.0;
  AffineInv a;
  a.A11 = M[4] * D; a.A22 = M[0] * D; a.A12 = -M[1] * D; a.A21 = -M[3] * D;
  a.b1 = -a.A11 * M[2] - a.A12 * M[5];
  a.b2 = -a.A21 * M[2] - a.A22 * M[5];
  warp_affine_kernel<<<dim3(ceil_div(W, 128), H), 128, 0, (cudaStream_t)stream>>>(src, H, W, C, a, dst);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

int bdrc_remap_bilinear_u8(bdrc_ctx* ctx, const uint8_t* src, int H, int W, int C, const float* coords,
                           uint8_t* dst, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  BDRC_REQUIRE(ctx, src && dst && coords && H > 0 && W > 0 && C >= 1 && C <= 4, "remap_bilinear_u8: bad arguments");
  remap_kernel<<<dim3(ceil_div(W, 128), H), 128, 0, (cudaStream_t)stream>>>(src, H, W, C, coords, dst);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

static int tps_remap_any(bdrc_ctx* ctx, bool cubic, const void* src, int H, int W, int C, const double* ctrl_host,
                        const double* params_host, int n_ctrl, uint8_t* dst, void* workspace, size_t workspace_bytes,
                        void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  BDRC_REQUIRE(ctx, src && dst && ctrl_host && params_host && n_ctrl >= 1 && n_ctrl <= TPS_MAX_CTRL && C >= 1 && C <= 4 &&
                   H > 0 && W > 0,
               "tps_remap: bad arguments (at most %d control points)", TPS_MAX_CTRL);
  BDRC_REQUIRE(ctx, workspace && workspace_bytes >= sizeof(TpsParams), "tps_remap: workspace too small");
  TpsParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.n = n_ctrl;
  for (int j = 0; j < n_ctrl; ++j) { tp.ctrl[j][0] = ctrl_host[2 * j]; tp.ctrl[j][1] = ctrl_host[2 * j + 1]; }
  for (int j = 0; j < n_ctrl + 3; ++j) { tp.par[j][0] = params_host[2 * j]; tp.par[j][1] = params_host[2 * j + 1]; }
  cudaStream_t s = (cudaStream_t)stream;
  BDRC_CUDA(ctx, cudaMemcpyAsync(workspace, &tp, sizeof(tp), cudaMemcpyHostToDevice, s));
  BDRC_CUDA(ctx, cudaStreamSynchronize(s));   // tp lives on this stack frame
  const dim3 grid(ceil_div(W, 128), H);
  if (cubic)
    tps_remap_kernel<true><<<grid, 128, 0, s>>>(src, H, W, C, reinterpret_cast<const TpsParams*>(workspace), dst);
  else
    tps_remap_kernel<false><<<grid, 128, 0, s>>>(src, H, W, C, reinterpret_cast<const TpsParams*>(workspace), dst);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

int bdrc_tps_remap_u8(bdrc_ctx* ctx, const uint8_t* src, int H, int W, int C, const double* ctrl_host,
                      const double* params_host, int n_ctrl, uint8_t* dst, void* workspace, size_t workspace_bytes,
                      void* stream) {
  return tps_remap_any(ctx, false, src, H, W, C, ctrl_host, params_host, n_ctrl, dst, workspace, workspace_bytes, stream);
}

int bdrc_tps_remap_cubic_u8(bdrc_ctx* ctx, const double* coef, int H, int W, int C, const double* ctrl_host,
                            const double* params_host, int n_ctrl, uint8_t* dst, void* workspace, size_t workspace_bytes,
                            void* stream) {
  return tps_remap_any(ctx, true, coef, H, W, C, ctrl_host, params_host, n_ctrl, dst, workspace, workspace_bytes, stream);
}

size_t bdrc_spline3_coef_bytes(int H, int W, int C) {
  return H > 0 && W > 0 && C > 0 ? (size_t)H * W * C * sizeof(double) : 0;
}

size_t bdrc_spline3_workspace_bytes(int H, int W, int C) { return bdrc_spline3_coef_bytes(H, W, C); }

int bdrc_spline3_prefilter_u8(bdrc_ctx* ctx, const uint8_t* src, int H, int W, int C, double* coef, void* workspace,
                              size_t workspace_bytes, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  BDRC_REQUIRE(ctx, src && coef && H > 0 && W > 0 && C >= 1 && C <= 4, "spline3_prefilter_u8: bad arguments");
  BDRC_REQUIRE(ctx, workspace && workspace_bytes >= bdrc_spline3_workspace_bytes(H, W, C),
               "spline3_prefilter_u8: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  double* tmp = reinterpret_cast<double*>(workspace);
  // axis 0 (down the columns): line l = x * C + ch, neighbouring lines are adjacent in memory
  const long long t0 = (long long)W * C * ceil_div(H, SPL_SEG);
  spline3_filter_kernel<uint8_t><<<(unsigned)((t0 + SPL_THREADS - 1) / SPL_THREADS), SPL_THREADS, 0, s>>>(
      src, tmp, W * C, H, W * C, 0, (size_t)W * C);
  BDRC_CHECK_LAUNCH(ctx);
  // axis 1 (along the rows): line l = y * C + ch
  const long long t1 = (long long)H * C * ceil_div(W, SPL_SEG);
  spline3_filter_kernel<double><<<(unsigned)((t1 + SPL_THREADS - 1) / SPL_THREADS), SPL_THREADS, 0, s>>>(
      tmp, coef, H * C, W, C, (size_t)W * C, (size_t)C);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

int bdrc_remap_cubic_u8(bdrc_ctx* ctx, const double* coef, int H, int W, int C, const double* coords, uint8_t* dst,
                        void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  BDRC_REQUIRE(ctx, coef && dst && coords && H > 0 && W > 0 && C >= 1 && C <= 4, "remap_cubic_u8: bad arguments");
  remap_cubic_kernel<<<dim3(ceil_div(W, 128), H), 128, 0, (cudaStream_t)stream>>>(coef, H, W, C, coords, dst);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

size_t bdrc_tps_workspace_bytes(void) { return sizeof(TpsParams); }

}  // extern "C"
