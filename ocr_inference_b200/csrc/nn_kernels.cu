// Memory-bound network kernels between the implicit-GEMM convolutions (see nn_kernels.cuh).
// All activations are bf16 NHWC; arithmetic is fp32.
#include <cstdlib>

#include "nn_kernels.cuh"
#include "tc_ptx.cuh"

namespace {

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    f[2 * q] = __low2float(h[q]);
    f[2 * q + 1] = __high2float(h[q]);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int q = 0; q < 4; ++q) h[q] = __floats2bfloat162_rn(f[2 * q], f[2 * q + 1]);
  return v;
}

// ---------------------------------------------------------------------------------------------
// stem 7x7 s2 p3, Cout = 64.  CTA = 4 output rows x 64 output cols; thread = 4 px x 16 couts.
// ---------------------------------------------------------------------------------------------
constexpr int ST_TH = 4, ST_TW = 64;
constexpr int ST_IR = ST_TH * 2 + 5;    // 13 input rows
constexpr int ST_IC = ST_TW * 2 + 5;    // 133 input cols
constexpr int ST_ICS = 136;             // padded stride

template <int CIN, typename TIn>
__global__ void __launch_bounds__(256)
stem_kernel(const TIn* __restrict__ in, float in_div, int H, int W, const float* __restrict__ wp,
            const float* __restrict__ bias, bf16* __restrict__ out) {
  extern __shared__ __align__(16) float st_smem[];
  float* s_in = st_smem;                              // [ST_IR][CIN][ST_ICS]
  float* s_w = st_smem + ST_IR * CIN * ST_ICS;        // [49*CIN][64]
  const int Ho = H / 2, Wo = W / 2;
  const int n = blockIdx.z, oy0 = blockIdx.y * ST_TH, ox0 = blockIdx.x * ST_TW;
  for (int i = threadIdx.x; i < 49 * CIN * 64 / 4; i += 256)
    reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(wp) + i);
  const int iy0 = oy0 * 2 - 3, ix0 = ox0 * 2 - 3;
  for (int i = threadIdx.x; i < ST_IR * ST_IC; i += 256) {
    const int r = i / ST_IC, c = i - r * ST_IC;
    const int iy = iy0 + r, ix = ix0 + c;
    const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
#pragma unroll
    for (int ch = 0; ch < CIN; ++ch) {
      float v = 0.f;
      if (ok) v = (float)in[(((size_t)n * H + iy) * W + ix) * CIN + ch] / in_div;
      s_in[(r * CIN + ch) * ST_ICS + c] = v;
    }
  }
  __syncthreads();
  const int cg = threadIdx.x & 3, pg = threadIdx.x >> 2;
  const int row = pg >> 4, col4 = (pg & 15) * 4;
  float acc[4][16];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
#pragma unroll 1
  for (int ch = 0; ch < CIN; ++ch) {
#pragma unroll 1
    for (int r = 0; r < 7; ++r) {
      const float* ip = s_in + ((row * 2 + r) * CIN + ch) * ST_ICS + col4 * 2;
      float x[13];
#pragma unroll
      for (int k = 0; k < 13; ++k) x[k] = ip[k];
      const float* wr = s_w + ((ch * 7 + r) * 7) * 64 + cg * 16;
#pragma unroll
      for (int s = 0; s < 7; ++s) {
        float wv[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 t = *reinterpret_cast<const float4*>(wr + s * 64 + q * 4);
          wv[4 * q] = t.x; wv[4 * q + 1] = t.y; wv[4 * q + 2] = t.z; wv[4 * q + 3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(x[2 * i + s], wv[j], acc[i][j]);
      }
    }
  }
  const int oy = oy0 + row;
  if (oy >= Ho) return;
  float b[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) b[j] = __ldg(bias + cg * 16 + j);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ox = ox0 + col4 + i;
    if (ox >= Wo) continue;
    float f[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = fmaxf(acc[i][j] + b[j], 0.f);
    uint4* o = reinterpret_cast<uint4*>(out + (((size_t)n * Ho + oy) * Wo + ox) * 64 + cg * 16);
    o[0] = pack8(f);
    o[1] = pack8(f + 8);
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void maxpool_kernel(const bf16* __restrict__ in, int N, int H, int W, int C,
                               bf16* __restrict__ out) {
  const int Ho = H / 2, Wo = W / 2, C8 = C / 8;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)N * Ho * Wo * C8;
  if (idx >= total) return;
  const int c8 = (int)(idx % C8);
  size_t p = idx / C8;
  const int ox = (int)(p % Wo); p /= Wo;
  const int oy = (int)(p % Ho);
  const int n = (int)(p / Ho);
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
  for (int dy = -1; dy <= 1; ++dy) {
    const int iy = oy * 2 + dy;
    if (iy < 0 || iy >= H) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int ix = ox * 2 + dx;
      if (ix < 0 || ix >= W) continue;
      float f[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(in + (((size_t)n * H + iy) * W + ix) * C + c8 * 8)), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], f[j]);
    }
  }
  *reinterpret_cast<uint4*>(out + (((size_t)n * Ho + oy) * Wo + ox) * C + c8 * 8) = pack8(m);
}

// Depthwise 3x3: each thread owns 8 channels of one column and walks a vertical strip of DW_STRIP
// output rows.  The 72 weights of its channel group live in registers for the whole strip (the
// first version re-read them from L1 for every output: 18 of its 21 loads per output were weights).
// DIL1: the three live input rows stay in registers (3 loads per output instead of 9).
template <bool DIL1, int DW_STRIP>
__global__ void __launch_bounds__(256)
depthwise_kernel(const bf16* __restrict__ in, int in_ld, int in_coff, int N, int H, int W, int C, int dil,
                 const float* __restrict__ w9c, bf16* __restrict__ out, int out_ld, int out_coff) {
  const int C8 = C / 8;
  const int strips = (H + DW_STRIP - 1) / DW_STRIP;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)N * strips * W * C8;
  if (idx >= total) return;
  const int c8 = (int)(idx % C8);
  size_t p = idx / C8;
  const int x = (int)(p % W); p /= W;
  const int st = (int)(p % strips);
  const int n = (int)(p / strips);
  const bf16* base = in + (size_t)n * H * W * in_ld + in_coff + c8 * 8;
  bf16* obase = out + (size_t)n * H * W * out_ld + out_coff + c8 * 8;
  float w[9][8];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w9c + k * C + c8 * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w9c + k * C + c8 * 8 + 4));
    w[k][0] = w0.x; w[k][1] = w0.y; w[k][2] = w0.z; w[k][3] = w0.w;
    w[k][4] = w1.x; w[k][5] = w1.y; w[k][6] = w1.z; w[k][7] = w1.w;
  }
  auto load_px = [&](int iy, int ix) -> uint4 {
    return (iy >= 0 && iy < H && ix >= 0 && ix < W)
               ? __ldg(reinterpret_cast<const uint4*>(base + ((size_t)iy * W + ix) * in_ld))
               : make_uint4(0u, 0u, 0u, 0u);
  };
  auto fma_px = [&](const uint4& v, int k, float* acc) {
    float f[8];
    unpack8(v, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[j], w[k][j], acc[j]);
  };
  const int y0 = st * DW_STRIP;
  if (DIL1) {
    uint4 r0[3], r1[3], r2[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) { r0[s] = load_px(y0 - 1, x + s - 1); r1[s] = load_px(y0, x + s - 1); }
#pragma unroll 4
    for (int i = 0; i < DW_STRIP; ++i) {
      const int y = y0 + i;
      if (y >= H) break;
#pragma unroll
      for (int s = 0; s < 3; ++s) r2[s] = load_px(y + 1, x + s - 1);
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int s = 0; s < 3; ++s) { fma_px(r0[s], s, acc); fma_px(r1[s], 3 + s, acc); fma_px(r2[s], 6 + s, acc); }
      *reinterpret_cast<uint4*>(obase + ((size_t)y * W + x) * out_ld) = pack8(acc);
#pragma unroll
      for (int s = 0; s < 3; ++s) { r0[s] = r1[s]; r1[s] = r2[s]; }
    }
  } else {
#pragma unroll 2
    for (int i = 0; i < DW_STRIP; ++i) {
      const int y = y0 + i;
      if (y >= H) break;
      uint4 v[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) v[r * 3 + s] = load_px(y + (r - 1) * dil, x + (s - 1) * dil);
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
      for (int k = 0; k < 9; ++k) fma_px(v[k], k, acc);
      *reinterpret_cast<uint4*>(obase + ((size_t)y * W + x) * out_ld) = pack8(acc);
    }
  }
}

// Depthwise 3x3, dilation 1, TMA-fed row ring.  A CTA owns (image, 64-channel chunk, DWR_PX-pixel column
// band, segment of DWR_ROWS output rows).  Input rows ((DWR_PX + 2) px x 64 ch, halo and channel tail
// zero-filled by TMA out-of-bounds handling) stream through an 8-slot shared-memory ring, each loaded
// exactly once; a thread reads the neighbours of its pixels in the arriving row and feeds the three output
// rows that row contributes to (no vertical re-reads).  Summation order per output is (r, s) row-major.
constexpr int DWR_PX = 32, DWR_SLOTS = 8, DWR_ROWS = 32;
constexpr int DWR_SLOT_BYTES = (DWR_PX + 2) * 128;

// Thread mapping: lane = channel PAIR (32 lanes x 2 = the 64 channels of the chunk), warp = 4 consecutive pixels:
// 18 register-resident weights per thread (a (pixel, 8 channels) mapping needs 72 and fits only 2 CTAs per SM),
// conflict-free 128-byte shared-memory rows, neighbouring pixels of a thread share their loads (6 LDS per 4
// outputs), a warp's store is one full 128-byte line.  ncu: the kernel is instruction-issue bound (69 % of the
// issue slots), not bandwidth bound; packed fma.rn.f32x2 made it slower.
__global__ void __launch_bounds__(256)
depthwise_ring_kernel(const __grid_constant__ CUtensorMap tmap, int N, int H, int W, int C,
                       const float* __restrict__ w9c, bf16* __restrict__ out, int out_ld, int out_coff, int* errflag) {
  extern __shared__ __align__(128) uint8_t dwr_smem_raw[];
  uint8_t* ring = dwr_smem_raw + ((128u - (tcptx::smem_u32(dwr_smem_raw) & 127u)) & 127u);
  __shared__ __align__(8) uint64_t full_bar[DWR_SLOTS];
  const int xt_n = W / DWR_PX, chunks = (C + 63) / 64, segs = (H + DWR_ROWS - 1) / DWR_ROWS;
  int item = blockIdx.x;
  const int xt = item % xt_n; item /= xt_n;
  const int seg = item % segs; item /= segs;
  const int ch = item % chunks;
  const int n = item / chunks;
  const int x0 = xt * DWR_PX, c0 = ch * 64;
  const int ya = seg * DWR_ROWS, yb = min(H, ya + DWR_ROWS);
  const int rows_in = yb - ya + 2;             // input rows ya-1 .. yb
  const int tid = threadIdx.x;
  const int lane = tid & 31, pb = (tid >> 5) * 4;
  const bool cvalid = c0 + 2 * lane < C;

  if (tid == 0) {
    for (int i = 0; i < DWR_SLOTS; ++i) tcptx::mbar_init(tcptx::smem_u32(&full_bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap)) : "memory");
  }
  __syncthreads();
  auto issue = [&](int k) {   // input row ya - 1 + k -> slot k % SLOTS
    const uint32_t bar = tcptx::smem_u32(&full_bar[k % DWR_SLOTS]);
    tcptx::mbar_arrive_expect_tx(bar, DWR_SLOT_BYTES);
    tcptx::tma_load_4d(tcptx::smem_u32(ring + (size_t)(k % DWR_SLOTS) * DWR_SLOT_BYTES), &tmap, bar, c0, x0 - 1,
                       ya - 1 + k, n);
  };
  if (tid == 0)
    for (int k = 0; k < DWR_SLOTS && k < rows_in; ++k) issue(k);

  float2 w[9];
#pragma unroll
  for (int k = 0; k < 9; ++k)
    w[k] = cvalid ? __ldg(reinterpret_cast<const float2*>(w9c + k * C + c0 + 2 * lane)) : make_float2(0.f, 0.f);
  float2 a_prev[4], a_cur[4], a_next[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) a_prev[q] = a_cur[q] = a_next[q] = make_float2(0.f, 0.f);
  bf16* obase = out + ((size_t)n * H * W + x0 + pb) * out_ld + out_coff + c0 + 2 * lane;

  for (int k = 0; k < rows_in; ++k) {
    const int i = ya - 1 + k;                  // input row
    tcptx::mbar_wait(tcptx::smem_u32(&full_bar[k % DWR_SLOTS]), (uint32_t)((k / DWR_SLOTS) & 1), errflag, 9);
    const uint8_t* row = ring + (size_t)(k % DWR_SLOTS) * DWR_SLOT_BYTES + pb * 128 + lane * 4;
    uint32_t v[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) v[j] = *reinterpret_cast<const uint32_t*>(row + j * 128);
    // The slot is refilled by TMA (async proxy) right after the barrier.  Arriving at a barrier does not wait
    // for this thread's outstanding shared-memory LOADS (the scoreboard only stalls an instruction that reads
    // their destination registers) and an async-proxy write is not ordered behind them, so a refill could
    // overtake a load still in flight: wrong pixels, different from run to run (tools/dw_check.py,
    // tests/test_gpu_determinism.py).  __threadfence_block() compiles to MEMBAR.SC.CTA + DEPBAR on the load
    // scoreboard: every thread's loads have returned before it arrives.  ("Using" the registers in an asm
    // statement does not help: ptxas drops instructions whose result is unused.)
    __threadfence_block();
    __syncthreads();
    if (tid == 0 && k + DWR_SLOTS < rows_in) issue(k + DWR_SLOTS);
    float2 f[6];   // bf16 -> fp32 is a 16-bit shift: low half << 16, high half masked
#pragma unroll
    for (int j = 0; j < 6; ++j) f[j] = make_float2(__uint_as_float(v[j] << 16), __uint_as_float(v[j] & 0xFFFF0000u));
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        // (packed fma.rn.f32x2 / FFMA2 was tried here: 0.93 ms against 0.85 ms per 148 tiles with scalar FFMA)
        a_next[q].x = fmaf(f[q + s].x, w[s].x, a_next[q].x);          // r = 0 row of output i + 1
        a_next[q].y = fmaf(f[q + s].y, w[s].y, a_next[q].y);
        a_cur[q].x = fmaf(f[q + s].x, w[3 + s].x, a_cur[q].x);        // r = 1 row of output i
        a_cur[q].y = fmaf(f[q + s].y, w[3 + s].y, a_cur[q].y);
        a_prev[q].x = fmaf(f[q + s].x, w[6 + s].x, a_prev[q].x);      // r = 2 row of output i - 1
        a_prev[q].y = fmaf(f[q + s].y, w[6 + s].y, a_prev[q].y);
      }
      if (i - 1 >= ya && cvalid) {
        const __nv_bfloat162 o = __floats2bfloat162_rn(a_prev[q].x, a_prev[q].y);
        *reinterpret_cast<__nv_bfloat162*>(obase + ((size_t)(i - 1) * W + q) * out_ld) = o;
      }
      a_prev[q] = a_cur[q]; a_cur[q] = a_next[q]; a_next[q] = make_float2(0.f, 0.f);
    }
  }
}

// The three dilated depthwise 3x3 branches of the ASPP (rates 12 / 24 / 36 on a 32x32 map, ipynb:L411-L420)
// in ONE launch: a CTA pulls the whole 32x32 map of one (image, 32-channel chunk) into shared memory
// with a single TMA box (64 KB) and produces the three branch outputs from it, so the input is read
// once instead of 3 x (up to) 9 times and two launches disappear.  Taps that fall outside the map are
// skipped (at rate 36 only the centre tap survives).  Summation order per output is (r, s) row-major.
// 32-channel chunks, not 64: two CTAs fit an SM, so one computes while the other waits for its map (with one
// 128 KB CTA per SM the load and the 8-warp compute phase alternated: 0.25 ms per 148 tiles against 0.095 ms of HBM time).
constexpr int ASPP_HW = 32;
constexpr int ASPP_CH = 32;                       // channels per CTA
constexpr int ASPP_PXB = ASPP_CH * 2;             // bytes per pixel in the slab
constexpr int ASPP_GROUPS = 256 / (ASPP_CH / 8);  // pixel groups (threads per 8-channel group)
__global__ void __launch_bounds__(256, 2)
aspp_dw3_kernel(const __grid_constant__ CUtensorMap tmap, int C, int r0, int r1, int r2,
                const float* __restrict__ w0, const float* __restrict__ w1, const float* __restrict__ w2,
                bf16* __restrict__ o0, bf16* __restrict__ o1, bf16* __restrict__ o2, float* __restrict__ mean_out,
                int* errflag) {
  extern __shared__ __align__(128) uint8_t aspp_smem_raw[];
  uint8_t* slab = aspp_smem_raw + ((128u - (tcptx::smem_u32(aspp_smem_raw) & 127u)) & 127u);
  __shared__ __align__(8) uint64_t bar;
  __shared__ float part[ASPP_GROUPS][ASPP_CH];
  const int chunks = C / ASPP_CH;
  const int ch = blockIdx.x % chunks, n = blockIdx.x / chunks;
  const int tid = threadIdx.x, c8 = tid % (ASPP_CH / 8), g = tid / (ASPP_CH / 8);
  if (tid == 0) {
    tcptx::mbar_init(tcptx::smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    tcptx::mbar_arrive_expect_tx(tcptx::smem_u32(&bar), ASPP_HW * ASPP_HW * ASPP_PXB);
    tcptx::tma_load_4d(tcptx::smem_u32(slab), &tmap, tcptx::smem_u32(&bar), ch * ASPP_CH, 0, 0, n);
  }
  tcptx::mbar_wait(tcptx::smem_u32(&bar), 0, errflag, 10);
  constexpr int ITERS = ASPP_HW * ASPP_HW / ASPP_GROUPS;
  if (mean_out) {
    // global average of the map (the ASPP image-pooling branch, ipynb:L421) while it is in shared memory;
    // fixed summation order: ITERS pixels per thread, then the thread groups in order
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    for (int j = 0; j < ITERS; ++j) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(slab + (size_t)(g + ASPP_GROUPS * j) * ASPP_PXB + c8 * 16), f);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] += f[q];
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) part[g][c8 * 8 + q] = acc[q];
    __syncthreads();
    if (tid < ASPP_CH) {
      float sum = 0.f;
      for (int k = 0; k < ASPP_GROUPS; ++k) sum += part[k][tid];
      mean_out[(size_t)n * C + ch * ASPP_CH + tid] = sum / (float)(ASPP_HW * ASPP_HW);
    }
  }
  const int rates[3] = {r0, r1, r2};
  const float* ws[3] = {w0, w1, w2};
  bf16* outs[3] = {o0, o1, o2};
#pragma unroll 1
  for (int b = 0; b < 3; ++b) {
    const int d = rates[b];
    // channel pairs as packed fp32x2 operands (each half an IEEE fmaf, same order of taps): half the FMA instructions
    ulonglong2 w[9][2];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      w[k][0] = __ldg(reinterpret_cast<const ulonglong2*>(ws[b] + k * C + ch * ASPP_CH + c8 * 8));
      w[k][1] = __ldg(reinterpret_cast<const ulonglong2*>(ws[b] + k * C + ch * ASPP_CH + c8 * 8 + 4));
    }
    bf16* ob = outs[b] + (size_t)n * ASPP_HW * ASPP_HW * C + ch * ASPP_CH + c8 * 8;
#pragma unroll 2
    for (int j = 0; j < ITERS; ++j) {
      const int p = g + ASPP_GROUPS * j;      // a warp covers 8 consecutive pixels: 512 contiguous bytes of the slab
      const int y = p / ASPP_HW, x = p % ASPP_HW;
      uint64_t acc2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int iy = y + (r - 1) * d;
        if (iy < 0 || iy >= ASPP_HW) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int ix = x + (s - 1) * d;
          if (ix < 0 || ix >= ASPP_HW) continue;
          const uint4 v = *reinterpret_cast<const uint4*>(slab + (size_t)(iy * ASPP_HW + ix) * ASPP_PXB + c8 * 16);
          const uint32_t u[4] = {v.x, v.y, v.z, v.w};
          const ulonglong2 w0 = w[r * 3 + s][0], w1 = w[r * 3 + s][1];
          const uint64_t wk[4] = {w0.x, w0.y, w1.x, w1.y};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint64_t f2;
            asm("mov.b64 %0, {%1, %2};" : "=l"(f2) : "r"(u[q] << 16), "r"(u[q] & 0xFFFF0000u));
            asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[q]) : "l"(f2), "l"(wk[q]));
          }
        }
      }
      float acc[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * q]), "=f"(acc[2 * q + 1]) : "l"(acc2[q]));
      *reinterpret_cast<uint4*>(ob + (size_t)p * C) = pack8(acc);
    }
  }
}

// x4 (or xf) align_corners=True upsample, strip variant: each thread produces UP_STRIP consecutive
// output rows of one column for 8 channels and re-loads a low-res row pair only when it changes.
// grid = (columns x channel groups / 256, strips, images): everything that depends on the output row -- the
// source rows, the vertical weights, whether a low-res row has to be (re)loaded -- is uniform over the block, so
// it lives in uniform registers and the row-change tests are uniform branches (ncu r01c: this kernel used 77 % of
// the issue slots with ~117 instructions per 16-byte store when the row was derived from a flat thread index).
constexpr int UP_STRIP = 8;
__global__ void __launch_bounds__(256)
upsample_ac_strip_kernel(const bf16* __restrict__ in, int H, int W, int C, int f, bf16* __restrict__ out,
                         int out_ld, int out_coff) {
  const int Ho = H * f, Wo = W * f, C8 = C / 8;
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= Wo * C8) return;
  const int x = e / C8, c8 = e - x * C8;
  const int st = blockIdx.y, n = blockIdx.z;
  const float sy = (float)(H - 1) / (float)(Ho - 1), sx = (float)(W - 1) / (float)(Wo - 1);
  const float fx = sx * (float)x;
  const int x0 = (int)fx, x1 = x0 + (x0 < W - 1 ? 1 : 0);
  const float lx1 = fx - (float)x0, lx0 = 1.f - lx1;
  const bf16* b0 = in + ((size_t)n * H * W + x0) * C + c8 * 8;
  const bf16* b1 = in + ((size_t)n * H * W + x1) * C + c8 * 8;
  bf16* op = out + (((size_t)n * Ho + st * UP_STRIP) * Wo + x) * out_ld + out_coff + c8 * 8;
  const size_t orow = (size_t)Wo * out_ld;
  int cy0 = -1, cy1 = -1;
  // both interpolations as packed fp32x2 operations on channel pairs (each half an IEEE mul / fma): the kernel is bound
  // by instruction issue, and the packed forms halve the 32 multiply-adds per stored 16 bytes
  auto pk = [](float lo, float hi) { uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; };
  auto mul2 = [](uint64_t x, uint64_t y) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(x), "l"(y)); return d; };
  auto fma2 = [](uint64_t x, uint64_t y, uint64_t z) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(y), "l"(z));
    return d;
  };
  const uint64_t LX0 = pk(lx0, lx0), LX1 = pk(lx1, lx1);
  uint64_t a0[4], a1[4];   // horizontally interpolated low-res rows cy0 / cy1
  auto hrow = [&](int yy, uint64_t* dst) {
    float v0[8], v1[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(b0 + (size_t)yy * W * C)), v0);
    unpack8(__ldg(reinterpret_cast<const uint4*>(b1 + (size_t)yy * W * C)), v1);
#pragma unroll
    for (int j = 0; j < 4; ++j) dst[j] = fma2(LX1, pk(v1[2 * j], v1[2 * j + 1]), mul2(LX0, pk(v0[2 * j], v0[2 * j + 1])));
  };
#pragma unroll
  for (int i = 0; i < UP_STRIP; ++i) {
    const int y = st * UP_STRIP + i;
    if (y >= Ho) break;
    const float fy = sy * (float)y;
    const int y0 = (int)fy, y1 = y0 + (y0 < H - 1 ? 1 : 0);
    const float ly1 = fy - (float)y0, ly0 = 1.f - ly1;
    if (y0 != cy0) {
      if (y0 == cy1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) a0[j] = a1[j];
      } else {
        hrow(y0, a0);
      }
      cy0 = y0;
    }
    if (y1 != cy1) {
      if (y1 == cy0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) a1[j] = a0[j];
      } else {
        hrow(y1, a1);
      }
      cy1 = y1;
    }
    const uint64_t LY0 = pk(ly0, ly0), LY1 = pk(ly1, ly1);
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t r = fma2(LY1, a1[j], mul2(LY0, a0[j]));
      float lo, hi;
      asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
      const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
      w[j] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(op + (size_t)i * orow) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// align_corners=True bilinear, fp32 math in torch's operand order (UpSampleBilinear2d)
__global__ void upsample_ac_kernel(const bf16* __restrict__ in, int N, int H, int W, int C, int f,
                                   bf16* __restrict__ out, int out_ld, int out_coff) {
  const int Ho = H * f, Wo = W * f, C8 = C / 8;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)N * Ho * Wo * C8;
  if (idx >= total) return;
  const int c8 = (int)(idx % C8);
  size_t p = idx / C8;
  const int x = (int)(p % Wo); p /= Wo;
  const int y = (int)(p % Ho);
  const int n = (int)(p / Ho);
  const float sy = (float)(H - 1) / (float)(Ho - 1), sx = (float)(W - 1) / (float)(Wo - 1);
  const float fy = sy * (float)y, fx = sx * (float)x;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
  const float ly1 = fy - (float)y0, lx1 = fx - (float)x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const bf16* b = in + (size_t)n * H * W * C + c8 * 8;
  float v00[8], v01[8], v10[8], v11[8], r[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(b + ((size_t)y0 * W + x0) * C)), v00);
  unpack8(__ldg(reinterpret_cast<const uint4*>(b + ((size_t)y0 * W + x1) * C)), v01);
  unpack8(__ldg(reinterpret_cast<const uint4*>(b + ((size_t)y1 * W + x0) * C)), v10);
  unpack8(__ldg(reinterpret_cast<const uint4*>(b + ((size_t)y1 * W + x1) * C)), v11);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    r[j] = ly0 * (lx0 * v00[j] + lx1 * v01[j]) + ly1 * (lx0 * v10[j] + lx1 * v11[j]);
  *reinterpret_cast<uint4*>(out + (((size_t)n * Ho + y) * Wo + x) * out_ld + out_coff + c8 * 8) = pack8(r);
}

__global__ void upsample_logits_kernel(const float* __restrict__ in, int N, int h, int w, int C,
                                       float* __restrict__ out) {
  const int Ho = h * 4, Wo = w * 4;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)N * C * Ho * Wo;
  if (idx >= total) return;
  size_t p = idx;
  const int x = (int)(p % Wo); p /= Wo;
  const int y = (int)(p % Ho); p /= Ho;
  const int c = (int)(p % C);
  const int n = (int)(p / C);
  const float sy = (float)(h - 1) / (float)(Ho - 1), sx = (float)(w - 1) / (float)(Wo - 1);
  const float fy = sy * (float)y, fx = sx * (float)x;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly1 = fy - (float)y0, lx1 = fx - (float)x0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
  const float* b = in + (size_t)n * h * w * C + c;
  out[idx] = bdrc_bilerp(b[((size_t)y0 * w + x0) * C], b[((size_t)y0 * w + x1) * C], b[((size_t)y1 * w + x0) * C],
                         b[((size_t)y1 * w + x1) * C], lx0, lx1, ly0, ly1);
}

// one CTA per image: mean over HW -> 1x1 conv + ReLU -> tail of the projection conv
__global__ void __launch_bounds__(512)
aspp_pool_bias_kernel(const bf16* __restrict__ x, int HW, int Cin, const float* __restrict__ Wp,
                      const float* __restrict__ bp, int Cmid, const float* __restrict__ Wtail,
                      const float* __restrict__ bproj, int Cout, int Cout_pad,
                      float* __restrict__ bias_eff, const float* __restrict__ mean_in) {
  extern __shared__ float pb_smem[];
  float* g = pb_smem;          // [Cin]
  float* h = pb_smem + Cin;    // [Cmid]
  float* part = h + Cmid;      // [slices][Cin] partial sums
  const int n = blockIdx.x;
  const bf16* xb = x + (size_t)n * HW * Cin;
  if (mean_in) {
    for (int c = threadIdx.x; c < Cin; c += blockDim.x) g[c] = mean_in[(size_t)n * Cin + c];
  } else {
  // mean over HW: thread = (8-channel group, pixel slice); 128-bit loads, fixed-order combine
    const int C8 = Cin / 8;
    const int slices = blockDim.x / C8 > 0 ? blockDim.x / C8 : 1;
    const int cg = threadIdx.x % C8, ps = threadIdx.x / C8;
    if (ps < slices) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      for (int p = ps; p < HW; p += slices) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(xb + (size_t)p * Cin + cg * 8)), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += f[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) part[(size_t)ps * Cin + cg * 8 + j] = acc[j];
    }
    __syncthreads();
    // fixed summation order -> bit-reproducible from run to run
    for (int c = threadIdx.x; c < Cin; c += blockDim.x) {
      float s = 0.f;
      for (int q = 0; q < slices; ++q) s += part[(size_t)q * Cin + c];
      g[c] = s / (float)HW;
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int o = warp; o < Cmid; o += nw) {
    float s = 0.f;
    for (int c = lane; c < Cin; c += 32) s = fmaf(Wp[(size_t)o * Cin + c], g[c], s);
    for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    if (lane == 0) h[o] = fmaxf(s + bp[o], 0.f);
  }
  __syncthreads();
  for (int o = warp; o < Cout_pad; o += nw) {
    float s = 0.f;
    if (o < Cout) {
      for (int c = lane; c < Cmid; c += 32) s = fmaf(Wtail[(size_t)o * Cmid + c], h[c], s);
      for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
      s += bproj[o];
    }
    if (lane == 0) bias_eff[(size_t)n * Cout_pad + o] = s;
  }
}

// Squeeze-excite as two short, fully parallel launches:
//   se_partial_gate_kernel : L * SE_CH CTAs; CTA (l, q) sums chunk q of line l's d column-wise (fixed order) into
//                            partial[l][q][C]; the LAST CTA of a line to finish (a ticket counter) adds the SE_CH
//                            partials in chunk order -- bit-reproducible whichever CTA it is -- and evaluates the
//                            gate MLP once per line, weights staged in shared memory with every load in flight;
//   se_apply_kernel        : data = relu(res + d * gate), elementwise, d from L2 where the conv just left it.
// History (B200, 64 lines, per unit): one-CTA-per-line gate kernel + apply 21 + 16 us; a thread-block-cluster
// version with the partial sums in distributed shared memory 37-54 us (clusters of 4-8 CTAs with a 32-64 KB slab
// schedule in too few waves); per-chunk CTAs that each evaluate the MLP 50 us (19 M warp instructions, every CTA
// pays the weight fetch latency); this version: see profiles/README.md.
constexpr int SE_CH = 8;
__global__ void __launch_bounds__(256)
se_partial_gate_kernel(const bf16* __restrict__ d, int T, int C, float* __restrict__ partial, int* __restrict__ tickets,
                       const float* __restrict__ W1, const float* __restrict__ b1, int Cmid, const float* __restrict__ W2,
                       const float* __restrict__ b2, float* __restrict__ gate, int direct) {
  extern __shared__ float se_sm[];                 // [slices][C] partial sums, later the MLP's staging
  __shared__ int last;
  // programmatic dependent launch (se_fused_launch, pdl = 1): the next kernel of the chain may be scheduled as soon as
  // every CTA of this one is resident; nothing is read or written before the previous kernel has completed
  tcptx::griddep_launch();
  tcptx::griddep_wait();
  const int q = blockIdx.x % SE_CH, l = blockIdx.x / SE_CH;
  const int Tq = (T + SE_CH - 1) / SE_CH;
  const int t0 = q * Tq, t1 = min(T, t0 + Tq);
  const bf16* db = d + (size_t)l * T * C;
  const int C8 = C / 8;
  const int slices = blockDim.x / C8 > 0 ? blockDim.x / C8 : 1;
  const int cg = threadIdx.x % C8, ps = threadIdx.x / C8;
  if (ps < slices) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int t = t0 + ps; t < t1; t += 4 * slices) {      // four rows in flight per thread, summed in time order
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int tt = t + u * slices;
        v[u] = tt < t1 ? __ldg(reinterpret_cast<const uint4*>(db + (size_t)tt * C + cg * 8)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (t + u * slices < t1) {
          float f[8];
          unpack8(v[u], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) se_sm[(size_t)ps * C + cg * 8 + j] = acc[j];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float sum = 0.f;
    for (int k = 0; k < slices; ++k) sum += se_sm[(size_t)k * C + c];
    partial[((size_t)l * SE_CH + q) * C + c] = sum;
  }
  __threadfence();                                  // this CTA's partial is visible before its ticket is drawn
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(&tickets[l], 1) == SE_CH - 1);
  __syncthreads();
  if (!last) return;
  __threadfence();
  // ---- the line's gate, once: mean -> FC -> ReLU -> FC -> sigmoid ----
  float* m = se_sm;                              // [C]
  float* z = m + C;                              // [Cmid]
  float* w1s = z + Cmid;                         // [Cmid][C]
  float* w2s = w1s + (size_t)Cmid * C;           // [C][Cmid + 1]  (padded: a thread per row reads conflict-free)
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < SE_CH; ++k) sum += __ldcg(partial + ((size_t)l * SE_CH + k) * C + c);
    m[c] = sum / (float)T;
  }
  if (direct) {
    // The MLP straight from the weights in L2 (same summation order as the staged form below, so the same bits):
    // a warp per hidden unit reads its W1 row in 128-byte pieces, a thread per channel its own W2 row as float4s,
    // every load of a round in flight together -- two L2 round trips instead of a 64 KB staging pass.
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int o = warp; o < Cmid; o += nw) {
      const float* wr = W1 + (size_t)o * C;
      float sum = 0.f;
      if (C == 256) {
        float w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = __ldg(wr + lane + 32 * k);
#pragma unroll
        for (int k = 0; k < 8; ++k) sum = fmaf(w[k], m[lane + 32 * k], sum);
      } else {
        for (int c = lane; c < C; c += 32) sum = fmaf(__ldg(wr + c), m[c], sum);
      }
      for (int k = 16; k > 0; k >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, k);
      if (lane == 0) z[o] = fmaxf(sum + b1[o], 0.f);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < C; o += blockDim.x) {
      const float* wr = W2 + (size_t)o * Cmid;
      float sum = b2[o];
      if (Cmid == 32) {
        float4 w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = __ldg(reinterpret_cast<const float4*>(wr) + k);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          sum = fmaf(w[k].x, z[4 * k], sum);
          sum = fmaf(w[k].y, z[4 * k + 1], sum);
          sum = fmaf(w[k].z, z[4 * k + 2], sum);
          sum = fmaf(w[k].w, z[4 * k + 3], sum);
        }
      } else {
        for (int c = 0; c < Cmid; ++c) sum = fmaf(__ldg(wr + c), z[c], sum);
      }
      gate[(size_t)l * C + o] = 1.f / (1.f + expf(-sum));
    }
    if (threadIdx.x == 0) tickets[l] = 0;
    return;
  }
  for (int i = threadIdx.x; i < Cmid * C / 4; i += blockDim.x)
    reinterpret_cast<float4*>(w1s)[i] = __ldg(reinterpret_cast<const float4*>(W1) + i);
  for (int i = threadIdx.x; i < C * Cmid; i += blockDim.x) w2s[(i / Cmid) * (Cmid + 1) + i % Cmid] = __ldg(W2 + i);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int o = warp; o < Cmid; o += nw) {
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) sum = fmaf(w1s[(size_t)o * C + c], m[c], sum);
    for (int k = 16; k > 0; k >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, k);
    if (lane == 0) z[o] = fmaxf(sum + b1[o], 0.f);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < C; o += blockDim.x) {
    float sum = b2[o];
    for (int c = 0; c < Cmid; ++c) sum = fmaf(w2s[(size_t)o * (Cmid + 1) + c], z[c], sum);
    gate[(size_t)l * C + o] = 1.f / (1.f + expf(-sum));
  }
  if (threadIdx.x == 0) tickets[l] = 0;            // ready for the next unit / call
}

__global__ void __launch_bounds__(256)
se_apply_kernel(const bf16* __restrict__ res, int res_ld, const bf16* __restrict__ d, const float* __restrict__ gate, int T,
                int C, size_t total8, bf16* __restrict__ out, int out_ld) {
  const int C8 = C / 8;
  tcptx::griddep_launch();
  tcptx::griddep_wait();
  // two 16-byte items per thread, both loads in flight before either is used
  const size_t i0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  uint4 rv[2], dv4[2];
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const size_t idx = i0 + u;
    if (idx < total8) {
      const size_t row = idx / C8;
      rv[u] = __ldg(reinterpret_cast<const uint4*>(res + row * res_ld + (idx - row * C8) * 8));
      dv4[u] = __ldg(reinterpret_cast<const uint4*>(d) + idx);
    }
  }
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const size_t idx = i0 + u;
    if (idx < total8) {
      const size_t row = idx / C8;          // l * T + t
      const int c8 = (int)(idx - row * C8), l = (int)(row / T);
      float r[8], dv[8], o[8];
      unpack8(rv[u], r);
      unpack8(dv4[u], dv);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + (size_t)l * C + c8 * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + (size_t)l * C + c8 * 8 + 4));
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaf(dv[j], g[j], r[j]), 0.f);
      *reinterpret_cast<uint4*>(out + row * out_ld + c8 * 8) = pack8(o);
    }
  }
}

}  // namespace

int stem_conv_launch(bdrc_ctx* ctx, const void* in, int cin, float in_div, int N, int H, int W,
                     const float* w_packed, const float* bias, bf16* out, cudaStream_t s) {
  BDRC_REQUIRE(ctx, (cin == 1 || cin == 3) && H % 2 == 0 && W % 2 == 0, "stem: cin must be 1 or 3");
  dim3 grid(ceil_div(W / 2, ST_TW), ceil_div(H / 2, ST_TH), N);
  const size_t smem = (size_t)(ST_IR * cin * ST_ICS + 49 * cin * 64) * sizeof(float);
  if (cin == 1) {
    auto k = stem_kernel<1, uint8_t>;
    BDRC_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 256, smem, s>>>(reinterpret_cast<const uint8_t*>(in), in_div, H, W, w_packed, bias, out);
  } else {
    auto k = stem_kernel<3, float>;
    BDRC_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, 256, smem, s>>>(reinterpret_cast<const float*>(in), in_div, H, W, w_packed, bias, out);
  }
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

int maxpool3x3s2_launch(bdrc_ctx* ctx, const bf16* in, int N, int H, int W, int C, bf16* out,
                        cudaStream_t s) {
  BDRC_REQUIRE(ctx, C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool: C %% 8 != 0");
  const size_t total = (size_t)N * (H / 2) * (W / 2) * (C / 8);
  maxpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, N, H, W, C, out);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

int depthwise3x3_launch(bdrc_ctx* ctx, const bf16* in, int in_ld, int in_coff, int N, int H, int W,
                        int C, int dil, const float* w9c, bf16* out, int out_ld, int out_coff,
                        cudaStream_t s) {
  BDRC_REQUIRE(ctx, C % 8 == 0 && in_ld % 8 == 0 && in_coff % 8 == 0 && out_ld % 8 == 0 && out_coff % 8 == 0,
               "depthwise: channel counts must be multiples of 8");
  if (dil == 1 && W % DWR_PX == 0 && bdrc_tma_encode_fn() != nullptr) {
    CUtensorMap map;
    const cuuint64_t ld = (cuuint64_t)in_ld;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)(DWR_PX + 2), 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = bdrc_tma_encode_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(in) + in_coff, dims,
                                      strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return bdrc_fail(ctx, BDRC_ERR_CUDA, "depthwise: cuTensorMapEncodeTiled failed: %d", (int)r);
    const int items = N * ((C + 63) / 64) * ((H + DWR_ROWS - 1) / DWR_ROWS) * (W / DWR_PX);
    depthwise_ring_kernel<<<items, 256, DWR_SLOTS * DWR_SLOT_BYTES + 128, s>>>(map, N, H, W, C, w9c, out, out_ld,
                                                                              out_coff, ctx->d_errflag);
    BDRC_CHECK_LAUNCH(ctx);
    return 0;
  }
  if (dil == 1) {
    const size_t total = (size_t)N * ((H + 15) / 16) * W * (C / 8);
    depthwise_kernel<true, 16><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, in_ld, in_coff, N, H, W, C, dil, w9c,
                                                                              out, out_ld, out_coff);
  } else {
    const size_t total = (size_t)N * ((H + 3) / 4) * W * (C / 8);
    depthwise_kernel<false, 4><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, in_ld, in_coff, N, H, W, C, dil, w9c,
                                                                              out, out_ld, out_coff);
  }
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

int aspp_dw3_launch(bdrc_ctx* ctx, const bf16* x, int N, int C, const int rates[3], const float* const w9c[3],
                    bf16* const out[3], float* mean_out, cudaStream_t s) {
  BDRC_REQUIRE(ctx, C % ASPP_CH == 0 && bdrc_tma_encode_fn() != nullptr, "aspp_dw3: C %% 32 != 0 or no TMA encoder");
  CUtensorMap map;
  cuuint64_t dims[4] = {(cuuint64_t)C, ASPP_HW, ASPP_HW, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)ASPP_HW * C * 2, (cuuint64_t)ASPP_HW * ASPP_HW * C * 2};
  cuuint32_t box[4] = {ASPP_CH, ASPP_HW, ASPP_HW, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = bdrc_tma_encode_fn()(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(x), dims, strides, box,
                                    es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return bdrc_fail(ctx, BDRC_ERR_CUDA, "aspp_dw3: cuTensorMapEncodeTiled failed: %d", (int)r);
  const size_t smem = (size_t)ASPP_HW * ASPP_HW * ASPP_PXB + 128;
  static bool attr_set = false;
  if (!attr_set) {
    BDRC_CUDA(ctx, cudaFuncSetAttribute(aspp_dw3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  aspp_dw3_kernel<<<N * (C / ASPP_CH), 256, smem, s>>>(map, C, rates[0], rates[1], rates[2], w9c[0], w9c[1], w9c[2], out[0],
                                                 out[1], out[2], mean_out, ctx->d_errflag);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

int upsample_bilinear_ac_launch(bdrc_ctx* ctx, const bf16* in, int N, int H, int W, int C, int f,
                                bf16* out, int out_ld, int out_coff, cudaStream_t s) {
  BDRC_REQUIRE(ctx, C % 8 == 0 && out_ld % 8 == 0 && out_coff % 8 == 0, "upsample: C %% 8 != 0");
  const dim3 grid((unsigned)ceil_div(W * f * (C / 8), 256), (unsigned)ceil_div(H * f, UP_STRIP), (unsigned)N);
  BDRC_REQUIRE(ctx, grid.y <= 65535 && grid.z <= 65535, "upsample: too many strips / images for one launch");
  upsample_ac_strip_kernel<<<grid, 256, 0, s>>>(in, H, W, C, f, out, out_ld, out_coff);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

int upsample_logits_launch(bdrc_ctx* ctx, const float* in, int N, int h, int w, int C, float* out,
                           cudaStream_t s) {
  const size_t total = (size_t)N * C * h * 4 * w * 4;
  upsample_logits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(in, N, h, w, C, out);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

int aspp_pool_bias_launch(bdrc_ctx* ctx, const bf16* x, int N, int HW, int Cin, const float* Wp,
                          const float* bp, int Cmid, const float* Wtail, const float* bproj,
                          int Cout, int Cout_pad, float* bias_eff, const float* mean_in, cudaStream_t s) {
  aspp_pool_bias_kernel<<<N, 512, (size_t)(Cin + Cmid + (512 / (Cin / 8)) * Cin) * sizeof(float), s>>>(
      x, HW, Cin, Wp, bp, Cmid, Wtail, bproj, Cout, Cout_pad, bias_eff, mean_in);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

int se_fused_launch(bdrc_ctx* ctx, const bf16* res, int res_ld, const bf16* d, int L, int T, int C, const float* W1,
                    const float* b1, int Cmid, const float* W2, const float* b2, bf16* out, int out_ld, float* scratch,
                    int pdl, cudaStream_t s) {
  BDRC_REQUIRE(ctx, C % 8 == 0 && C / 8 <= 256 && res_ld % 8 == 0 && out_ld % 8 == 0 && Cmid >= 1 && scratch &&
                        (Cmid * C) % 4 == 0,
               "se_fused: C, res_ld, out_ld must be multiples of 8");
  // scratch (fp32 words, se_fused_scratch_floats(L, C)): [L][SE_CH][C] partial sums | [L][C] gate | [L] tickets (zeroed
  // once per forward by the caller with se_fused_reset)
  float* partial = scratch;
  float* gate = partial + (size_t)L * SE_CH * C;
  int* tickets = reinterpret_cast<int*>(gate + (size_t)L * C);
  const int slices = 256 / (C / 8) > 0 ? 256 / (C / 8) : 1;
  size_t smem = (size_t)slices * C * sizeof(float);
  const size_t mlp = (size_t)(C + Cmid + Cmid * C + C * (Cmid + 1)) * sizeof(float);
  if (mlp > smem) smem = mlp;
  BDRC_REQUIRE(ctx, smem <= 200 * 1024, "se_fused: gate MLP too large for shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    BDRC_CUDA(ctx, cudaFuncSetAttribute(se_partial_gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  // BDRC_SE_STAGED=1 keeps the round-2 form (weights staged in shared memory) for A/B timing; same results
  static const int direct = []() { const char* e = getenv("BDRC_SE_STAGED"); return (e && e[0] == '1') ? 0 : 1; }();
  const int dflag = direct && (reinterpret_cast<uintptr_t>(W2) & 15) == 0 ? 1 : 0;
  const size_t total8 = (size_t)L * T * (C / 8);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.stream = s;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  cfg.gridDim = dim3((unsigned)(L * SE_CH));
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  BDRC_CUDA(ctx, cudaLaunchKernelEx(&cfg, se_partial_gate_kernel, d, T, C, partial, tickets, W1, b1, Cmid, W2, b2, gate, dflag));
  cfg.gridDim = dim3((unsigned)((total8 + 511) / 512));
  cfg.dynamicSmemBytes = 0;
  BDRC_CUDA(ctx, cudaLaunchKernelEx(&cfg, se_apply_kernel, res, res_ld, d, (const float*)gate, T, C, total8, out, out_ld));
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

size_t se_fused_scratch_floats(int L, int C) { return (size_t)L * SE_CH * C + (size_t)L * C + (size_t)L + 16; }

int se_fused_reset(bdrc_ctx* ctx, float* scratch, int L, int C, cudaStream_t s) {
  BDRC_CUDA(ctx, cudaMemsetAsync(scratch + (size_t)L * SE_CH * C + (size_t)L * C, 0, (size_t)L * sizeof(int), s));
  return 0;
}

int se_fused_chunks(void) { return SE_CH; }

