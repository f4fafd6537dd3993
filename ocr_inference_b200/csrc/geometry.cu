// GPU connected-component labelling of the line mask: the pixel-scanning half of the host geometry
// (BDRC/line_detection.py:33-44, 362-421).
//
// What the reference computes with cv2 on the host, per page:
//   * cv2.boundingRect(mask) and, inside get_line_threshold (:362-421), ~70 findContours calls on 20-pixel-wide
//     vertical slices: the number of contours of each slice and, for the slice with the most, the bounding-box
//     centres of its contours -- a census whose only output is one float (median centre // count);
//   * the contours themselves (polygons: approxPolyDP, convex hulls, filled ROI masks), which stay cv2 work on the
//     host because their vertices depend on OpenCV's border following and polygon approximation.
//
// The census needs no polygon, only connectivity, so it is computed here exactly:
//   findContours(RETR_LIST) of a binary slice returns one outer border per 8-connected foreground component and
//   one hole border per hole, a hole being a 4-connected background component that does not reach the slice's
//   frame.  boundingRect of an outer border is the component's bounding box; boundingRect of a hole border is
//   the hole's bounding box grown by one pixel on every side (the border runs over the foreground pixels that
//   are 4-adjacent to the hole, and the pixel above the top-most / below the bottom-most hole pixel is always
//   one of them).
// Labelling: union-find with atomicMin over horizontal RUNS (a pixel-level union-find spends milliseconds collapsing
// the page-sized background component: measured 2.3 ms per page; run-level it is tens of microseconds): every pixel
// first points at the first pixel of its run (one block scan per row), runs of neighbouring rows are united once
// per overlap (N for both classes, NW / NE for the 8-connected foreground where no N link exists), run starts are
// flattened, and a pixel's root is its run start's root.  The
// slices are laid side by side in one "strip" image with an empty column between neighbours and an empty frame
// around everything (exactly the strip the host code builds, ocr_inference_b200/line_detection.py), so pixel 0
// is background and the outer background is the one component whose root is 0: every other background root is
// a hole.  All counters are integer atomics: results do not depend on scheduling; only the ORDER of the centre
// list does, and its consumer takes a median.
//
// Also here: bdrc_mask_components, the same labelling over a whole mask with per-component bounding box and
// pixel count (cv2.connectedComponentsWithStats, 8-connectivity), the "contour-lite" primitive.
#include <climits>

#include "common.cuh"

namespace {

constexpr int CENSUS_MAX_SLICES = 256;
constexpr int CENSUS_MAX_CENTERS = 4096;
constexpr int HDR_INTS = 16;
// header layout (int32): 0 x, 1 y, 2 w, 3 h (cv2.boundingRect of the non-zero pixels; w = 0: empty mask),
// 4 x_steps, 5 strip width (with frame), 6 strip height (with frame), 7 best slice, 8 its contour count,
// 9 centres written, 10 overflow flag (centres beyond CENSUS_MAX_CENTERS, or more slices than CENSUS_MAX_SLICES)

__global__ void occupancy_kernel(const uint8_t* __restrict__ mask, int H, int W, uint8_t* __restrict__ row_any,
                                 uint8_t* __restrict__ col_any) {
  const int y = blockIdx.x;
  const uint8_t* row = mask + (size_t)y * W;
  int any = 0;
  if ((W & 15) == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0) {
    for (int x16 = threadIdx.x; x16 < W / 16; x16 += blockDim.x) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(row) + x16);
      if (v.x | v.y | v.z | v.w) {
        any = 1;
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if ((w[q] >> (8 * j)) & 255u) col_any[x16 * 16 + q * 4 + j] = 1;   // every writer stores the same 1
      }
    }
  } else {
    for (int x = threadIdx.x; x < W; x += blockDim.x)
      if (row[x]) { any = 1; col_any[x] = 1; }
  }
  if (__syncthreads_or(any) && threadIdx.x == 0) row_any[y] = 1;
}

__global__ void __launch_bounds__(1024) bbox_kernel(const uint8_t* __restrict__ row_any, const uint8_t* __restrict__ col_any,
                                                    int H, int W, int slice_width, int* __restrict__ hdr,
                                                    int* __restrict__ counts) {
  __shared__ int s[4];
  if (threadIdx.x == 0) { s[0] = INT_MAX; s[1] = -1; s[2] = INT_MAX; s[3] = -1; }
  __syncthreads();
  int x0 = INT_MAX, x1 = -1, y0 = INT_MAX, y1 = -1;
  for (int x = threadIdx.x; x < W; x += blockDim.x)
    if (col_any[x]) { x0 = min(x0, x); x1 = max(x1, x); }
  for (int y = threadIdx.x; y < H; y += blockDim.x)
    if (row_any[y]) { y0 = min(y0, y); y1 = max(y1, y); }
  atomicMin(&s[0], x0); atomicMax(&s[1], x1); atomicMin(&s[2], y0); atomicMax(&s[3], y1);
  for (int i = threadIdx.x; i < CENSUS_MAX_SLICES; i += blockDim.x) counts[i] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < HDR_INTS; ++i) hdr[i] = 0;
    if (s[1] >= 0) {
      const int w = s[1] - s[0] + 1, h = s[3] - s[2] + 1;
      int steps = (w / slice_width) / 2;                        // :383-385
      hdr[0] = s[0]; hdr[1] = s[2]; hdr[2] = w; hdr[3] = h;
      if (steps > CENSUS_MAX_SLICES) { hdr[10] = 1; steps = 0; }
      hdr[4] = steps;
      hdr[5] = steps * (slice_width + 1) + 1;
      hdr[6] = h + 2;
    }
  }
}

// strip pixel (r, c), frame included.  Columns: 0 = frame, then per slice `sw` pixel columns + 1 separator.
__device__ __forceinline__ int strip_pixel(const uint8_t* __restrict__ mask, int W, const int* hdr, int sw, int r, int c) {
  const int SH = hdr[6];
  if (r == 0 || r == SH - 1 || c == 0) return 0;
  const int slice = (c - 1) / (sw + 1), k = (c - 1) - slice * (sw + 1);
  if (k == sw) return 0;
  const int x = hdr[0] + hdr[4] * (slice + 1) + k;             // x0 = x + x_steps * step (:389), step = slice + 1
  if (x >= W) return 0;
  return mask[(size_t)(hdr[1] + r - 1) * W + x] != 0;
}

__global__ void strip_init_kernel(const uint8_t* __restrict__ mask, int W, const int* __restrict__ hdr, int sw, int stride,
                                  uint8_t* __restrict__ pix, int* __restrict__ ymin, int* __restrict__ ymax) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= hdr[5] || r >= hdr[6]) return;
  const uint32_t p = (uint32_t)r * stride + c;
  pix[p] = (uint8_t)strip_pixel(mask, W, hdr, sw, r, c);
  ymin[p] = INT_MAX;
  ymax[p] = -1;
}

__device__ __forceinline__ uint32_t uf_find(volatile uint32_t* label, uint32_t x) {
  uint32_t l = label[x];
  while (l != x) { x = l; l = label[x]; }
  return x;
}

__device__ __forceinline__ void uf_union(uint32_t* label, uint32_t a, uint32_t b) {
  bool done;
  do {
    a = uf_find(label, a);
    b = uf_find(label, b);
    if (a < b) {
      const uint32_t old = atomicMin(&label[b], a);
      done = (old == b);
      b = old;
    } else if (b < a) {
      const uint32_t old = atomicMin(&label[a], b);
      done = (old == a);
      a = old;
    } else {
      done = true;
    }
  } while (!done);
}

// label[p] = index of the first pixel of p's horizontal run (pixels of one class).  One block per row.
__global__ void __launch_bounds__(256) run_label_kernel(const uint8_t* __restrict__ pix, uint32_t* __restrict__ label,
                                                        const int* __restrict__ dims, int Wfix, int Hfix, int stride) {
  __shared__ int carry[256];
  const int r = blockIdx.x;
  const int Wd = dims ? dims[5] : Wfix, Hd = dims ? dims[6] : Hfix;
  if (r >= Hd) return;
  const uint8_t* row = pix + (size_t)r * stride;
  const int per = (Wd + 255) / 256;
  const int c0 = min(Wd, (int)threadIdx.x * per), c1 = min(Wd, c0 + per);
  int last = -1;
  for (int c = c0; c < c1; ++c)
    if (c == 0 || row[c] != row[c - 1]) last = c;
  carry[threadIdx.x] = last;
  __syncthreads();
  for (int o = 1; o < 256; o <<= 1) {             // inclusive max-scan
    const int v = threadIdx.x >= o ? carry[threadIdx.x - o] : -1;
    __syncthreads();
    carry[threadIdx.x] = max(carry[threadIdx.x], v);
    __syncthreads();
  }
  int cur = threadIdx.x ? carry[threadIdx.x - 1] : -1;
  for (int c = c0; c < c1; ++c) {
    if (c == 0 || row[c] != row[c - 1]) cur = c;
    label[(size_t)r * stride + c] = (uint32_t)r * stride + cur;
  }
}

__device__ __forceinline__ bool is_run_start(const uint8_t* pix, uint32_t p, int c) { return c == 0 || pix[p] != pix[p - 1]; }
// the run start of pixel p: start pixels may already have been re-pointed by a union, other pixels never are
__device__ __forceinline__ uint32_t run_start(const uint8_t* pix, const uint32_t* label, uint32_t p, int c) {
  return is_run_start(pix, p, c) ? p : label[p];
}

// unites the runs of row r with the runs of row r - 1 they touch: once per overlap (at its left-most pixel)
__global__ void ccl_merge_kernel(const uint8_t* __restrict__ pix, uint32_t* label, const int* __restrict__ dims, int Wfix, int Hfix,
                                 int stride) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  const int Wd = dims ? dims[5] : Wfix, Hd = dims ? dims[6] : Hfix;
  if (c >= Wd || r >= Hd || r == 0) return;
  const uint32_t p = (uint32_t)r * stride + c, up = p - stride;
  const uint8_t v = pix[p];
  if (pix[up] == v) {
    const uint32_t sp = run_start(pix, label, p, c), su = run_start(pix, label, up, c);
    const int cs = (int)(sp - (uint32_t)r * stride), cu = (int)(su - (uint32_t)(r - 1) * stride);
    if (c == max(cs, cu)) uf_union(label, sp, su);
  } else if (v) {
    // 8-connectivity: a diagonal link matters only where neither N nor the horizontal neighbour makes it
    if (c > 0 && pix[up - 1] && !pix[p - 1]) uf_union(label, run_start(pix, label, p, c), run_start(pix, label, up - 1, c - 1));
    if (c + 1 < Wd && pix[up + 1] && !pix[p + 1]) uf_union(label, run_start(pix, label, p, c), up + 1);   // NE starts its run (N is background)
  }
}

__global__ void ccl_flatten_starts_kernel(const uint8_t* __restrict__ pix, uint32_t* label, const int* __restrict__ dims, int Wfix,
                                          int Hfix, int stride) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  const int Wd = dims ? dims[5] : Wfix, Hd = dims ? dims[6] : Hfix;
  if (c >= Wd || r >= Hd) return;
  const uint32_t p = (uint32_t)r * stride + c;
  if (is_run_start(pix, p, c)) label[p] = uf_find(label, p);
}

// after ccl_flatten_starts_kernel: the root of any pixel
__device__ __forceinline__ uint32_t root_of(const uint8_t* pix, const uint32_t* label, uint32_t p, int c) {
  return is_run_start(pix, p, c) ? label[p] : label[label[p]];
}

__global__ void census_stats_kernel(const uint8_t* __restrict__ pix, uint32_t* label, const int* __restrict__ hdr, int sw, int stride,
                                    int* ymin, int* ymax, int* counts) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= hdr[5] || r >= hdr[6]) return;
  const uint32_t p = (uint32_t)r * stride + c;
  const uint32_t root = root_of(pix, label, p, c);
  if (root == 0) return;                 // outer background
  atomicMin(&ymin[root], r);
  atomicMax(&ymax[root], r);
  if (root == p) atomicAdd(&counts[(c - 1) / (sw + 1)], 1);   // one contour per foreground component / hole
}

__global__ void __launch_bounds__(1024) census_pick_kernel(const uint8_t* __restrict__ pix, const uint32_t* __restrict__ label,
                                                           int* hdr, int sw, int stride, const int* __restrict__ ymin,
                                                           const int* __restrict__ ymax, const int* __restrict__ counts,
                                                           int* __restrict__ centers) {
  __shared__ int best, n_out;
  const int steps = hdr[4];
  if (threadIdx.x == 0) {
    int b = 0;
    for (int i = 1; i < steps; ++i)
      if (counts[i] > counts[b]) b = i;          // the first slice with the highest count (stable sort, :409-411)
    best = b;
    n_out = 0;
  }
  __syncthreads();
  if (steps == 0) return;
  const int SH = hdr[6];
  const int c0 = 1 + best * (sw + 1);
  for (int i = threadIdx.x; i < (SH - 2) * sw; i += blockDim.x) {
    const int r = 1 + i / sw, c = c0 + i % sw;
    const uint32_t p = (uint32_t)r * stride + c;
    if (label[p] != p || p == 0) continue;      // roots are run starts that point at themselves
    int cy, ch;
    if (pix[p]) { cy = ymin[p] - 1; ch = ymax[p] - ymin[p] + 1; }
    else        { cy = ymin[p] - 2; ch = ymax[p] - ymin[p] + 3; }   // hole border: the hole grown by one pixel
    const int k = atomicAdd(&n_out, 1);
    if (k < CENSUS_MAX_CENTERS) centers[k] = cy + ch / 2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    hdr[7] = best;
    hdr[8] = counts[best];
    hdr[9] = min(n_out, CENSUS_MAX_CENTERS);
    if (n_out > CENSUS_MAX_CENTERS) hdr[10] = 1;
  }
}

// ---- whole-mask components ------------------------------------------------------------------------------
__global__ void comp_init_kernel(const uint8_t* __restrict__ mask, int H, int W, int stride, uint8_t* __restrict__ pix,
                                 uint32_t* __restrict__ label) {
  // frame of one background pixel around the mask: pixel 0 is background, the outer background has root 0
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= W + 2 || r >= H + 2) return;
  const uint32_t p = (uint32_t)r * stride + c;
  pix[p] = (r >= 1 && r <= H && c >= 1 && c <= W) ? (mask[(size_t)(r - 1) * W + (c - 1)] != 0) : 0;
}

struct CompStat { int x0, y0, x1, y1, area, root; };

__global__ void comp_roots_kernel(const uint8_t* __restrict__ pix, uint32_t* label, int H, int W, int stride, int* n_comp,
                                  CompStat* stats, int max_comp, int* slot_of_root) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= W + 2 || r >= H + 2) return;
  const uint32_t p = (uint32_t)r * stride + c;
  if (label[p] == p && is_run_start(pix, p, c) && pix[p]) {
    const int k = atomicAdd(n_comp, 1);
    if (k < max_comp) {
      stats[k] = CompStat{INT_MAX, INT_MAX, -1, -1, 0, (int)p};
      slot_of_root[p] = k;
    } else {
      slot_of_root[p] = -1;
    }
  }
}

__global__ void comp_stats_kernel(const uint8_t* __restrict__ pix, const uint32_t* __restrict__ label, int H, int W, int stride,
                                  CompStat* stats, const int* __restrict__ slot_of_root) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c >= W + 2 || r >= H + 2) return;
  const uint32_t p = (uint32_t)r * stride + c;
  if (!pix[p]) return;
  const int k = slot_of_root[root_of(pix, label, p, c)];
  if (k < 0) return;
  atomicMin(&stats[k].x0, c - 1); atomicMax(&stats[k].x1, c - 1);
  atomicMin(&stats[k].y0, r - 1); atomicMax(&stats[k].y1, r - 1);
  atomicAdd(&stats[k].area, 1);
}

}  // namespace

extern "C" {

int bdrc_census_max_centers(void) { return CENSUS_MAX_CENTERS; }

size_t bdrc_mask_census_workspace_bytes(int H, int W, int slice_width) {
  if (H <= 0 || W <= 0 || slice_width <= 0) return 0;
  const size_t stride = (size_t)((W / slice_width) / 2) * (slice_width + 1) + 1 + 3;
  const size_t px = stride * (size_t)(H + 2);
  return align_up((size_t)H + W, 256) + align_up(CENSUS_MAX_SLICES * sizeof(int), 256) + align_up(px, 256) +
         3 * align_up(px * sizeof(int), 256) + 1024;
}

// mask u8 [H][W] on the device -> row_any u8 [H], header int32[16], centres int32[bdrc_census_max_centers()]
// (device pointers; the caller copies what it needs to the host with the mask)
int bdrc_mask_census(bdrc_ctx* ctx, const uint8_t* mask, int H, int W, int slice_width, uint8_t* row_any, int32_t* header,
                     int32_t* centers, void* workspace, size_t workspace_bytes, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  BDRC_REQUIRE(ctx, mask && row_any && header && centers && H > 0 && W > 0 && slice_width > 0 && slice_width <= 64,
               "mask_census: bad arguments");
  BDRC_REQUIRE(ctx, workspace && workspace_bytes >= bdrc_mask_census_workspace_bytes(H, W, slice_width),
               "mask_census: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int max_steps = (W / slice_width) / 2;
  const int stride = max_steps * (slice_width + 1) + 1 + 3;
  const size_t px = (size_t)stride * (H + 2);
  BDRC_REQUIRE(ctx, px < 0x7fffffffULL, "mask_census: mask too large");
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  uint8_t* col_any = w;                        w += align_up((size_t)H + W, 256);
  int* counts = reinterpret_cast<int*>(w);     w += align_up(CENSUS_MAX_SLICES * sizeof(int), 256);
  uint8_t* pix = w;                            w += align_up(px, 256);
  uint32_t* label = reinterpret_cast<uint32_t*>(w); w += align_up(px * sizeof(int), 256);
  int* ymin = reinterpret_cast<int*>(w);       w += align_up(px * sizeof(int), 256);
  int* ymax = reinterpret_cast<int*>(w);
  BDRC_CUDA(ctx, cudaMemsetAsync(row_any, 0, (size_t)H, s));
  BDRC_CUDA(ctx, cudaMemsetAsync(col_any, 0, (size_t)W, s));
  bdrc_prof_begin(ctx, "a-10.mask_census", 0, s);
  occupancy_kernel<<<H, 256, 0, s>>>(mask, H, W, row_any, col_any);
  bbox_kernel<<<1, 1024, 0, s>>>(row_any, col_any, H, W, slice_width, header, counts);
  if (max_steps > 0) {
    dim3 grid(ceil_div(stride, 128), H + 2);
    strip_init_kernel<<<grid, 128, 0, s>>>(mask, W, header, slice_width, stride, pix, ymin, ymax);
    run_label_kernel<<<H + 2, 256, 0, s>>>(pix, label, header, 0, 0, stride);
    ccl_merge_kernel<<<grid, 128, 0, s>>>(pix, label, header, 0, 0, stride);
    ccl_flatten_starts_kernel<<<grid, 128, 0, s>>>(pix, label, header, 0, 0, stride);
    census_stats_kernel<<<grid, 128, 0, s>>>(pix, label, header, slice_width, stride, ymin, ymax, counts);
    census_pick_kernel<<<1, 1024, 0, s>>>(pix, label, header, slice_width, stride, ymin, ymax, counts, centers);
    ctx->launches += 6;
  }
  bdrc_prof_end(ctx, s);
  ctx->launches += 1;
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

size_t bdrc_mask_components_workspace_bytes(int H, int W) {
  if (H <= 0 || W <= 0) return 0;
  const size_t px = (size_t)(W + 2 + 2) * (H + 2);
  return align_up(px, 256) + 2 * align_up(px * sizeof(int), 256) + 1024;
}

// 8-connected foreground components of mask u8 [H][W]: n_comp (device int) and up to max_comp records
// {x0, y0, x1, y1 (inclusive), area, root} in unspecified order (cv2.connectedComponentsWithStats without label 0)
int bdrc_mask_components(bdrc_ctx* ctx, const uint8_t* mask, int H, int W, int32_t* n_comp, int32_t* stats, int max_comp,
                         void* workspace, size_t workspace_bytes, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  BDRC_REQUIRE(ctx, mask && n_comp && stats && H > 0 && W > 0 && max_comp > 0, "mask_components: bad arguments");
  BDRC_REQUIRE(ctx, workspace && workspace_bytes >= bdrc_mask_components_workspace_bytes(H, W),
               "mask_components: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  const int stride = W + 2 + 2;
  const size_t px = (size_t)stride * (H + 2);
  BDRC_REQUIRE(ctx, px < 0x7fffffffULL, "mask_components: mask too large");
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  uint8_t* pix = w;                                   w += align_up(px, 256);
  uint32_t* label = reinterpret_cast<uint32_t*>(w);   w += align_up(px * sizeof(int), 256);
  int* slot = reinterpret_cast<int*>(w);
  BDRC_CUDA(ctx, cudaMemsetAsync(n_comp, 0, sizeof(int), s));
  dim3 grid(ceil_div(W + 2, 128), H + 2);
  bdrc_prof_begin(ctx, "a-10.mask_components", 0, s);
  comp_init_kernel<<<grid, 128, 0, s>>>(mask, H, W, stride, pix, label);
  run_label_kernel<<<H + 2, 256, 0, s>>>(pix, label, nullptr, W + 2, H + 2, stride);
  ccl_merge_kernel<<<grid, 128, 0, s>>>(pix, label, nullptr, W + 2, H + 2, stride);
  ccl_flatten_starts_kernel<<<grid, 128, 0, s>>>(pix, label, nullptr, W + 2, H + 2, stride);
  comp_roots_kernel<<<grid, 128, 0, s>>>(pix, label, H, W, stride, n_comp, reinterpret_cast<CompStat*>(stats), max_comp, slot);
  comp_stats_kernel<<<grid, 128, 0, s>>>(pix, label, H, W, stride, reinterpret_cast<CompStat*>(stats), slot);
  bdrc_prof_end(ctx, s);
  ctx->launches += 5;
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

}  // extern "C"
