// Stem of the segmentation network on the tensor cores, fused with the max-pool that follows it:
//   conv 7x7 stride 2 pad 3 (1 folded input channel -> 64) + bias + ReLU + maxpool 3x3 stride 2 pad 1
//   (PyTorch-LineInference.ipynb:L233-L238), u8 tiles [N][512][512] -> bf16 NHWC [N][128][128][64].
//
// The three input channels of a binarised tile are identical (BDRC/utils.py:796-815 replicates the
// mask), so the 3x49 stem weights fold to 49 per output channel and the conv is a GEMM with
// K = 7 rows x 8 (7 taps + one zero column) = 56, padded to 64 = one 128-byte swizzle row.
//
// One CTA produces a 4 x 16 patch of POOLED outputs: it needs the 9 x 33 stem pixels around it
//   1. stage the 23 x 80 byte input patch in shared memory (zero outside the image),
//   2. im2col: 297 (-> 384) A rows x 64 bf16, written by the threads in the SWIZZLE_128B K-major layout
//      (u8 -> bf16 is exact; the window of a stem pixel is 7 runs of 8 consecutive bytes),
//   3. 3 M-tiles x 4 tcgen05.mma (M128 N64 K16), fp32 accumulators in TMEM,
//   4. epilogue: TMEM -> +bias, ReLU, bf16 -> shared memory (aliasing the A tiles), zero for stem
//      pixels outside the 256 x 256 stem image (exact: every pooled window holds a valid pixel >= 0),
//   5. 3x3/s2 max over shared memory -> 16-byte stores of the pooled NHWC tile.
// The 403 MB stem activation of a 48-tile chunk never exists in HBM.
// The five phases of a patch are separated by block-wide barriers, so ONE resident CTA leaves the SM idle in
// every hand-over; a CTA is therefore sized (3 M-tiles = 256 TMEM columns, 60 KB of shared memory, <= 64
// registers) so that TWO are resident per SM and one's im2col / pooling overlaps the other's MMAs / epilogue
// (round 1: 4 x 32 patches, 5 M-tiles, 512 TMEM columns, one CTA per SM: 0.60 ms per 148 tiles).
#include "nn_kernels.cuh"
#include "tc_ptx.cuh"

using namespace tcptx;

namespace {

constexpr int SP = 512;                 // input tile size
constexpr int PT_H = 4, PT_W = 16;      // pooled outputs per CTA tile
constexpr int SR_H = 2 * PT_H + 1;      // 9 stem rows
constexpr int SR_W = 2 * PT_W + 1;      // 33 stem cols
constexpr int M_VALID = SR_H * SR_W;    // 297
constexpr int M_TILES = 3;              // 384 rows
constexpr uint32_t STEM_TMEM_COLS = 256;  // 3 x 64 accumulator columns, power of two; two CTAs share the SM's 512
constexpr int PATCH_ROWS = 4 * PT_H + 7;     // 23 input rows
constexpr int PATCH_PITCH = 4 * PT_W + 16;   // input columns per patch row; aligned words starting at image col 4*px0 - 8
constexpr int P1_PITCH = 2 * PATCH_PITCH;    // bytes per row of the bf16 copy of the patch
constexpr int STEM_THREADS = 512;
constexpr uint32_t A_TILE_BYTES = 128 * 128;  // 16 KB
constexpr uint32_t SMEM_A = M_TILES * A_TILE_BYTES;          // 48 KB (also the pooled staging S)
constexpr uint32_t SMEM_B = 64 * 128;                        // 8 KB
constexpr uint32_t SMEM_PATCH = PATCH_ROWS * P1_PITCH;       // 3680 B
constexpr uint32_t SMEM_TOTAL = 1024 + SMEM_A + SMEM_B + 4096 + 256;

__device__ __forceinline__ uint32_t u8pair_to_bf16x2(uint32_t word, int k) {
  // bytes k and k+1 of `word` -> two exact bf16 (0..255)
  const float f0 = __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7650 | k)) - 8388608.0f;
  const float f1 = __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7650 | (k + 1))) - 8388608.0f;
  return __byte_perm(__float_as_uint(f0), __float_as_uint(f1), 0x7632);
}

static_assert(SMEM_PATCH <= 3712, "bias lives at patch + 3712");
static_assert(M_TILES * 64 <= STEM_TMEM_COLS, "accumulators must fit the TMEM allocation");

__global__ void __launch_bounds__(STEM_THREADS, 2)
stem_tc_pool_kernel(const uint8_t* __restrict__ in, int N, const bf16* __restrict__ wb, const float* __restrict__ bias,
                    bf16* __restrict__ out, int* errflag) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tcptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                         // 5 x [128][128 B] swizzled; reused as S[640][128 B]
  uint8_t* sB = smem + SMEM_A;                // [64][128 B] swizzled
  uint8_t* sP = sB + SMEM_B;                  // input patch
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(sP + 4096);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
  float* sBias = reinterpret_cast<float*>(sP + 3712);   // 64 floats after the patch (3680 -> 3712)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time set-up: weights into the swizzled B tile, bias, barrier, TMEM -----------------
  for (int i = tid; i < 64 * 8; i += STEM_THREADS) {
    const int o = i >> 3, c = i & 7;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(wb + o * 64 + c * 8));
    *reinterpret_cast<uint4*>(sB + o * 128 + ((c ^ (o & 7)) << 4)) = v;
  }
  if (tid < 64) sBias[tid] = __ldg(bias + tid);
  if (tid == 0) {
    mbar_init(smem_u32(mma_bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(STEM_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // K columns 56..63 of every A row are never written by the im2col.  Their weights are zero, so they
  // only have to be finite: zero them once (later tiles leave finite bf16 activations there)
  for (int m = tid; m < M_TILES * 128; m += STEM_THREADS)
    *reinterpret_cast<uint4*>(sA + (size_t)m * 128 + ((7 ^ (m & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = make_idesc_bf16(128, 64);

  const int tiles_x = 128 / PT_W, tiles_y = 128 / PT_H;
  const int total = N * tiles_y * tiles_x;
  uint32_t phase = 0;

  // ---- per-thread constants of the tile loop (nothing below depends on the tile) ---------------
  // patch loader: thread i < 460 owns the aligned word wq of patch row pr_ (and the word after it)
  const bool p_live = tid < PATCH_ROWS * (PATCH_PITCH / 4);
  const int pr_ = tid / (PATCH_PITCH / 4), pwq = tid - pr_ * (PATCH_PITCH / 4);
  // im2col items i = tid + 512 * it -> (kernel row r, A row m): source offset in the bf16 patch (12 bits) and
  // destination offset in the swizzled A tiles (16 bits); 0xFFFFFFFF = no item / row beyond the 297 valid ones
  constexpr int IM_ITEMS = 7 * M_TILES * 128;
  constexpr int IM_IT = (IM_ITEMS + STEM_THREADS - 1) / STEM_THREADS;   // 6
  uint32_t im_item[IM_IT];
#pragma unroll
  for (int it = 0; it < IM_IT; ++it) {
    const int i = tid + it * STEM_THREADS;
    const int r = i / (M_TILES * 128), m = i - r * (M_TILES * 128);
    uint32_t pk = 0xFFFFFFFFu;
    if (i < IM_ITEMS && m < M_VALID) {
      const int ry = m / SR_W, rx = m - ry * SR_W;
      // the window of stem pixel (ry, rx), kernel row r: patch bytes 3 + 2 rx .. + 7 of row 2 ry + r; the bf16 copy
      // holds patch byte e + 1 at element e, so the window starts at byte 4 + 4 rx of its row: four aligned words
      const uint32_t src = (uint32_t)((2 * ry + r) * P1_PITCH + 4 + 4 * rx);
      const uint32_t dst = (uint32_t)(m * 128 + ((r ^ (m & 7)) << 4));
      pk = src | (dst << 12);
    }
    im_item[it] = pk;
  }
  // epilogue rows m = t * 128 + ml: bit t = row is one of the 297, bit 4 + t = stem row 0 of the patch, bit 8 + t = stem
  // column 0 of the patch (outside the image for the patches on the top / left edge)
  const int ml = (warp & 3) * 32 + lane;
  uint32_t ep_flags = 0;
#pragma unroll
  for (int t = 0; t < M_TILES; ++t) {
    const int m = t * 128 + ml;
    const int ry = m / SR_W, rx = m - ry * SR_W;
    if (m < M_VALID) ep_flags |= 1u << t;
    if (ry == 0) ep_flags |= 16u << t;
    if (rx == 0) ep_flags |= 256u << t;
  }

  auto load_patch = [&](int tile, uint32_t& lo, uint32_t& hi) {
    lo = 0u; hi = 0u;
    if (tile < total && p_live) {
      const int tx = tile % tiles_x;
      const int ty = (tile / tiles_x) % tiles_y;
      const int n = tile / (tiles_x * tiles_y);
      const int iy = 4 * (ty * PT_H) - 5 + pr_, ix = 4 * (tx * PT_W) - 8 + pwq * 4;
      if (iy >= 0 && iy < SP) {
        const uint8_t* rowp = in + ((size_t)n * SP + iy) * SP;
        if (ix >= 0 && ix < SP) lo = __ldg(reinterpret_cast<const uint32_t*>(rowp + ix));
        if (ix + 4 >= 0 && ix + 4 < SP) hi = __ldg(reinterpret_cast<const uint32_t*>(rowp + ix + 4));
      }
    }
  };
  uint32_t p_lo, p_hi;
  load_patch(blockIdx.x, p_lo, p_hi);

  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int tx = tile % tiles_x;
    const int ty = (tile / tiles_x) % tiles_y;
    const int n = tile / (tiles_x * tiles_y);
    const int px0 = tx * PT_W, py0 = ty * PT_H;

    // ---- 1. input patch (fetched one tile ahead) as bf16, shifted by one column: element e = patch byte e + 1 -----
    if (p_live) {
      const uint32_t w = __funnelshift_r(p_lo, p_hi, 8);
      *reinterpret_cast<uint2*>(sP + pr_ * P1_PITCH + pwq * 8) = make_uint2(u8pair_to_bf16x2(w, 0), u8pair_to_bf16x2(w, 2));
    }
    __syncthreads();   // also orders the previous tile's pooling reads of S before the rewrite of A
    load_patch(tile + gridDim.x, p_lo, p_hi);   // consumed at the top of the next iteration: the L2 latency hides behind this tile

    // ---- 2. im2col: chunk r of A row m = 8 consecutive bf16 of the patch copy (the 8th meets a zero weight) ----------
#pragma unroll
    for (int it = 0; it < IM_IT; ++it) {
      const uint32_t pk = im_item[it];
      if (pk != 0xFFFFFFFFu) {
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(sP + (pk & 0xFFFu));
        *reinterpret_cast<uint4*>(sA + (pk >> 12)) = make_uint4(wp[0], wp[1], wp[2], wp[3]);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();

    // ---- 3. MMAs ---------------------------------------------------------------------------------
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint64_t bdesc = make_sdesc(smem_u32(sB));
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
          for (int t = 0; t < M_TILES; ++t)   // tile-inner order: back-to-back MMAs hit different accumulators
            tc_mma_bf16(tmem_base + (uint32_t)(t * 64), make_sdesc(smem_u32(sA) + t * A_TILE_BYTES) + (uint64_t)(2 * k),
                        bdesc + (uint64_t)(2 * k), idesc, k > 0 ? 1u : 0u);
        tc_commit(smem_u32(mma_bar));
      }
      __syncwarp();
    }
    mbar_wait(smem_u32(mma_bar), phase, errflag, 8);
    phase ^= 1;
    tc_fence_after();

    // ---- 4. epilogue: warp w -> TMEM lanes 32*(w%4).., column group q = w/4 of every tile ------------
    {
      const int q = warp >> 2;
      float b[16];
#pragma unroll
      for (int j = 0; j < 16; j += 4) {   // 16-byte broadcast loads: the kernel is bound by shared-memory wavefronts
        const float4 b4 = *reinterpret_cast<const float4*>(sBias + q * 16 + j);
        b[j] = b4.x; b[j + 1] = b4.y; b[j + 2] = b4.z; b[j + 3] = b4.w;
      }
      // rows that are not stem pixels of the image store zeros (exact for the max: every window holds a valid value >= 0)
      const uint32_t dead = (~ep_flags & 7u) | (py0 == 0 ? (ep_flags >> 4) & 7u : 0u) | (px0 == 0 ? (ep_flags >> 8) & 7u : 0u);
#pragma unroll
      for (int t = 0; t < M_TILES; ++t) {
        uint32_t v[16];
        tc_ld16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(t * 64 + q * 16), v);
        tc_wait_ld();
        const uint32_t keep = (dead >> t) & 1u ? 0u : 0xFFFFFFFFu;
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          // bias, ReLU and the bf16 rounding in one conversion: d = {hi: relu(a), lo: relu(b)}
          const float f0 = __uint_as_float(v[2 * j]) + b[2 * j], f1 = __uint_as_float(v[2 * j + 1]) + b[2 * j + 1];
          uint32_t h;
          asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(f1), "f"(f0));
          w[j] = h & keep;
        }
        uint8_t* rowp = sA + (size_t)(t * 128 + ml) * 128;
        *reinterpret_cast<uint4*>(rowp + (((2 * q) ^ (ml & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(rowp + (((2 * q + 1) ^ (ml & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
      }
    }
    tc_fence_before();
    __syncthreads();

    // ---- 5. 3x3 stride-2 max over S -> pooled NHWC ---------------------------------------------------
    for (int i = tid; i < PT_H * PT_W * 8; i += STEM_THREADS) {
      const int c8 = i & 7, pc = (i >> 3) % PT_W, pr = (i >> 3) / PT_W;
      __nv_bfloat162 mx[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) mx[j] = __floats2bfloat162_rn(0.f, 0.f);   // post-ReLU values are >= 0
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int m = (2 * pr + dy) * SR_W + 2 * pc + dx;
          const uint4 v = *reinterpret_cast<const uint4*>(sA + (size_t)m * 128 + ((c8 ^ (m & 7)) << 4));
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
          for (int j = 0; j < 4; ++j) mx[j] = __hmax2(mx[j], h[j]);
        }
      uint4 o;
      o.x = *reinterpret_cast<uint32_t*>(&mx[0]); o.y = *reinterpret_cast<uint32_t*>(&mx[1]);
      o.z = *reinterpret_cast<uint32_t*>(&mx[2]); o.w = *reinterpret_cast<uint32_t*>(&mx[3]);
      *reinterpret_cast<uint4*>(out + (((size_t)n * 128 + py0 + pr) * 128 + px0 + pc) * 64 + c8 * 8) = o;
    }
    // the next iteration's first __syncthreads orders these reads before A is rewritten
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(STEM_TMEM_COLS) : "memory");
  }
}

}  // namespace

int stem_tc_pool_launch(bdrc_ctx* ctx, const uint8_t* tiles, int N, const bf16* w_bf16, const float* bias, bf16* out,
                        cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    BDRC_CUDA(ctx, cudaFuncSetAttribute(stem_tc_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL));
    attr_set = true;
  }
  const int total = N * (128 / PT_H) * (128 / PT_W);
  const int grid = total < 2 * ctx->sm_count ? total : 2 * ctx->sm_count;   // two resident CTAs per SM
  stem_tc_pool_kernel<<<grid, STEM_THREADS, SMEM_TOTAL, s>>>(tiles, N, w_bf16, bias, out, ctx->d_errflag);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}
