// The 128x128-resolution tail of the DeepLabV3+ decoder as ONE kernel (PyTorch-LineInference.ipynb:L455-L466):
//
//     depthwise 3x3 (304 ch, pad 1)  ->  1x1 conv 304 -> 256 + bias + ReLU  ->  head 1x1 256 -> C + bias
//
// Unfused, these three layers write and re-read two page-sized activations (3.8 GB + 3.2 GB per 384-tile step) and
// ran at 4.1 ms of a 29.9 ms step, all of it HBM traffic.  Here the depthwise output never leaves the SM: SIMT
// warps compute it from a TMA-staged halo tile and write it, as bf16, straight into the SWIZZLE_128B A tile that
// tcgen05.mma consumes; the 256-channel result stays in TMEM, where the epilogue applies bias + ReLU, rounds to bf16
// (what the unfused path stores) and folds the head's per-pixel dot product.  Only the low-res logits reach HBM.
//
// Tile: 8 rows x 16 columns of output pixels (M = 128), K = 304 in five 64-channel chunks, N = 256.
// Roles (768 threads):
//   warp 0      TMA producer of the halo ring (4 slots): per chunk the (10 x 18 pixel) x 64-channel box of the concat
//               buffer (zero padding and the 304 -> 320 channel tail by out-of-bounds fill)
//   warp 3      TMA producer of the weight ring (2 slots): the 64 x 256 tile of the chunk
//   warp 1      MMA issuer (4 x tcgen05.mma M128 N256 K16 per chunk), commits free the A / weight slots
//   warp 2      TMEM allocation
//   warps 4-19  depthwise, two groups of eight: group g produces the chunks with (chunk counter & 1) == g into A slot g,
//               so two chunks are in production at a time (four depthwise warps per scheduler instead of two: the FMA
//               and shared-memory latencies of one chunk hide behind the other's); warp = output row of the tile, lane =
//               pair of channels; taps accumulate in the order r, s of the standalone depthwise kernel (bit-identical
//               bf16 results)
//   warps 20-23 epilogue: one TMEM lane (= pixel) per thread
// Weights, A and the accumulators are double-buffered, the halo ring is four deep (a two-deep ring left the
// depthwise warps waiting for TMA: 14.5 us per tile), persistent grid.
#include <cstring>
#include <type_traits>

#include "nn_kernels.cuh"
#include "tc_ptx.cuh"

using namespace tcptx;

namespace {

constexpr int TH = 8, TW = 16;                    // output tile
constexpr int HH = TH + 2, HW = TW + 2;           // halo
constexpr int KC = 64;                            // channels per chunk
constexpr uint32_t HALO_BYTES = HH * HW * 128;    // 23040
constexpr uint32_t B_BYTES = 256 * 128;           // 32768
constexpr uint32_t A_BYTES = 128 * 128;           // 16384
constexpr int DT_THREADS = 768;
constexpr int MAX_HEAD = 8;
constexpr int HS = 4;                             // halo ring depth

struct DecTailParams {
  int n_img, H, W, C, chunks;      // C = 304, chunks = 5
  int tiles_x, tiles_y, total_tiles;
  const float* dw_w;               // [9][C] fp32
  float* out;                      // [n][H][W][head_C] fp32
  int head_C;
  uint32_t idesc;
  int* errflag;
};

// The epilogue's constants travel in the kernel parameters (constant bank): every thread of the epilogue needs the SAME
// bias / head weight at the same time, so they are constant operands of the FADD / FFMA themselves.  As broadcast
// shared-memory loads (the first version) they were 128 wavefronts per warp and tile on the resource this kernel is
// bound by -- the depthwise loads, the MMA operand reads and the TMA writes all go through shared memory.
template <int HC>
struct TailConsts {
  float bias[256];                 // 1x1 conv 304 -> 256
  float head_b[HC];
  float head_w[HC][256];           // fp32 holding the bf16-rounded head weights
};

template <int HC>     // head classes: compile-time, the per-pixel dot products live in registers
__global__ void __launch_bounds__(DT_THREADS, 1)
dec_tail_kernel(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapW, const DecTailParams p,
                const __grid_constant__ TailConsts<HC> kc_) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sB = smem;                                   // 2 x 32 KB (1024-aligned, SWIZZLE_128B)
  uint8_t* sA = sB + 2 * B_BYTES;                       // 2 x 16 KB (1024-aligned, SWIZZLE_128B, written by SIMT)
  uint8_t* sH = sA + 2 * A_BYTES;                       // HS x 23040 B halo (plain rows of 128 B)
  float* sDW = reinterpret_cast<float*>(sH + HS * HALO_BYTES);     // [9][chunks * 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sDW + 9 * p.chunks * KC);
  uint64_t* halo_full = bars;          // [HS]
  uint64_t* halo_empty = bars + 4;     // [HS]
  uint64_t* b_full = bars + 8;         // [2]
  uint64_t* b_empty = bars + 10;       // [2]
  uint64_t* a_full = bars + 12;        // [2]
  uint64_t* a_empty = bars + 14;       // [2]
  uint64_t* tfull = bars + 16;         // [2]
  uint64_t* tempty = bars + 18;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int CP = p.chunks * KC;        // padded channel count (320)

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapX)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapW)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < HS; ++i) {
      mbar_init(smem_u32(&halo_full[i]), 1);
      mbar_init(smem_u32(&halo_empty[i]), 8);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&b_full[i]), 1);
      mbar_init(smem_u32(&b_empty[i]), 1);
      mbar_init(smem_u32(&a_full[i]), 8);
      mbar_init(smem_u32(&a_empty[i]), 1);
      mbar_init(smem_u32(&tfull[i]), 1);
      mbar_init(smem_u32(&tempty[i]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // depthwise weights (zero beyond C)
  for (int i = threadIdx.x; i < 9 * CP; i += DT_THREADS) {
    const int k = i / CP, c = i - k * CP;
    sDW[i] = c < p.C ? __ldg(p.dw_w + (size_t)k * p.C + c) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int first = blockIdx.x, step = gridDim.x;

  if (warp == 0) {
    // ===================== TMA producer: halo ring =====================
    uint32_t it = 0;     // chunk counter over all tiles of this CTA
    for (int tile = first; tile < p.total_tiles; tile += step) {
      const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, n = tile / (p.tiles_x * p.tiles_y);
      for (int kc = 0; kc < p.chunks; ++kc, ++it) {
        const uint32_t s = it % HS, ph = (it / HS) & 1u;
        mbar_wait_relaxed(smem_u32(&halo_empty[s]), ph ^ 1u, p.errflag, 21);
        if (elect_one()) {
          mbar_arrive_expect_tx(smem_u32(&halo_full[s]), HALO_BYTES);
          tma_load_4d(smem_u32(sH + s * HALO_BYTES), &tmapX, smem_u32(&halo_full[s]), kc * KC, tx * TW - 1, ty * TH - 1, n);
        }
        __syncwarp();
      }
    }
  } else if (warp == 3) {
    // ===================== TMA producer: weight ring =====================
    uint32_t it = 0;
    for (int tile = first; tile < p.total_tiles; tile += step) {
      for (int kc = 0; kc < p.chunks; ++kc, ++it) {
        const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
        mbar_wait_relaxed(smem_u32(&b_empty[s]), ph ^ 1u, p.errflag, 22);
        if (elect_one()) {
          mbar_arrive_expect_tx(smem_u32(&b_full[s]), B_BYTES);
          tma_load_2d(smem_u32(sB + s * B_BYTES), &tmapW, smem_u32(&b_full[s]), kc * KC, 0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    uint32_t it = 0, tcount = 0;
    for (int tile = first; tile < p.total_tiles; tile += step, ++tcount) {
      const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
      mbar_wait_relaxed(smem_u32(&tempty[acc]), aph ^ 1u, p.errflag, 23);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * 256u;
      for (int kc = 0; kc < p.chunks; ++kc, ++it) {
        const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
        mbar_wait_relaxed(smem_u32(&b_full[s]), ph, p.errflag, 24);
        mbar_wait_relaxed(smem_u32(&a_full[s]), ph, p.errflag, 25);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = make_sdesc(smem_u32(sA + s * A_BYTES));
          const uint64_t bdesc = make_sdesc(smem_u32(sB + s * B_BYTES));
#pragma unroll
          for (int k = 0; k < KC / 16; ++k)
            tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc, (kc > 0 || k > 0) ? 1u : 0u);
          tc_commit(smem_u32(&a_empty[s]));      // both slots are free once these MMAs have read them
          tc_commit(smem_u32(&b_empty[s]));
          if (kc == p.chunks - 1) tc_commit(smem_u32(&tfull[acc]));
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4 && warp < 20) {
    // ===================== depthwise producers of the A operand =====================
    // warp = output row of the tile (16 pixels), lane = one PAIR of channels of the chunk.  Every shared-memory access of
    // the warp is then one 128-byte wavefront (a halo pixel's 64 channels; a row of the A tile) and a halo value is
    // loaded once per filter row: 54 + 9 loads per thread and chunk.  This kernel is bound by shared-memory wavefronts
    // (depthwise loads + MMA operand reads + TMA writes); with lane = (4 adjacent pixels, 8 channels) and 16-byte
    // loads the same work cost 144 load wavefronts per warp and chunk instead of 72 (216 before the weight loads'
    // bank conflicts were removed: 0.98 -> 0.84 ms per 148 tiles).
    const int row = (warp - 4) & 7;
    const uint32_t group = (uint32_t)(warp - 4) >> 3;
    uint32_t it = 0;
    for (int tile = first; tile < p.total_tiles; tile += step) {
      for (int kc = 0; kc < p.chunks; ++kc, ++it) {
        if ((it & 1u) != group) continue;
        const uint32_t s = it & 1u, ph = (it >> 1) & 1u;          // A slot
        const uint32_t hs = it % HS, hph = (it / HS) & 1u;        // halo slot
        mbar_wait_relaxed(smem_u32(&halo_full[hs]), hph, p.errflag, 26);
        const uint8_t* halo = sH + hs * HALO_BYTES + lane * 4;
        // the two channels of a lane are the two halves of a packed fp32x2 FMA (each half an IEEE fmaf): this kernel is
        // bound by instruction issue (ncu: 74 % of the issue slots, FMA pipe 39 %), and the packed form halves the 288
        // FMA instructions per thread and chunk
        uint64_t acc[TW];
#pragma unroll
        for (int q = 0; q < TW; ++q) acc[q] = 0ull;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          uint64_t w[3];
#pragma unroll
          for (int t = 0; t < 3; ++t) w[t] = *reinterpret_cast<const uint64_t*>(sDW + (r * 3 + t) * CP + kc * KC + lane * 2);
          const uint8_t* hrow = halo + (size_t)((row + r) * HW) * 128;
#pragma unroll
          for (int c = 0; c < HW; ++c) {          // halo column c feeds outputs q = c - t (tap t): taps in the order r, s
            const uint32_t u = *reinterpret_cast<const uint32_t*>(hrow + c * 128);
            uint64_t f;
            asm("mov.b64 %0, {%1, %2};" : "=l"(f) : "r"(u << 16), "r"(u & 0xFFFF0000u));
#pragma unroll
            for (int t = 0; t < 3; ++t) {
              const int q = c - t;
              if (q >= 0 && q < TW) asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[q]) : "l"(f), "l"(w[t]));
            }
          }
        }
        // the halo slot goes back to TMA: this thread's shared-memory loads must have landed (see nn_kernels.cu)
        __threadfence_block();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&halo_empty[hs]));
        mbar_wait_relaxed(smem_u32(&a_empty[s]), ph ^ 1u, p.errflag, 27);   // the MMAs that read this A slot two chunks ago
        // A tile: row m = pixel, 16-byte chunk (lane / 4) XOR-swizzled with (m & 7); the warp writes one 128-byte row
        uint8_t* a = sA + s * A_BYTES + (size_t)(row * TW) * 128 + (lane & 3) * 4;
#pragma unroll
        for (int q = 0; q < TW; ++q) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float((uint32_t)acc[q]), __uint_as_float((uint32_t)(acc[q] >> 32)));
          *reinterpret_cast<uint32_t*>(a + q * 128 + (((lane >> 2) ^ (q & 7)) << 4)) = *reinterpret_cast<const uint32_t*>(&h);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&a_full[s]));
      }
    }
  } else if (warp >= 20) {
    // ===================== epilogue: bias + ReLU + bf16 rounding + head =====================
    const int q4 = warp & 3;                      // TMEM lane quarter
    const int m = q4 * 32 + lane;                 // pixel of the tile
    uint32_t tcount = 0;
    for (int tile = first; tile < p.total_tiles; tile += step, ++tcount) {
      const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, n = tile / (p.tiles_x * p.tiles_y);
      const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
      mbar_wait_relaxed(smem_u32(&tfull[acc]), aph, p.errflag, 28);
      tc_fence_after();
      float hacc[HC];
#pragma unroll
      for (int c = 0; c < HC; ++c) hacc[c] = 0.f;
      const uint32_t tbase = tmem_base + ((uint32_t)(q4 * 32) << 16) + acc * 256u;
      // 16 accumulator columns: bias + ReLU + the bf16 rounding of the unfused path's stored activation in one packed
      // conversion, then the head's dot products; C0 is a compile-time constant (constant-bank operands)
      auto fold = [&](const uint32_t* v, auto c0_tag) {
        constexpr int C0 = decltype(c0_tag)::value;
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          uint32_t h;
          asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(h)
              : "f"(__uint_as_float(v[j + 1]) + kc_.bias[C0 + j + 1]), "f"(__uint_as_float(v[j]) + kc_.bias[C0 + j]));
          const float f0 = __uint_as_float(h << 16), f1 = __uint_as_float(h & 0xFFFF0000u);
#pragma unroll
          for (int c = 0; c < HC; ++c) {
            hacc[c] = fmaf(f0, kc_.head_w[c][C0 + j], hacc[c]);
            hacc[c] = fmaf(f1, kc_.head_w[c][C0 + j + 1], hacc[c]);
          }
        }
      };
      // the TMEM load of the next 16 columns is in flight while these are folded
      uint32_t va[16], vb[16];
      tc_ld16(tbase, va);
      auto step32 = [&](auto c0_tag) {
        constexpr int C0 = decltype(c0_tag)::value;
        tc_wait_ld();
        tc_ld16(tbase + (uint32_t)(C0 + 16), vb);
        fold(va, std::integral_constant<int, C0>{});
        tc_wait_ld();
        if (C0 + 32 < 256) tc_ld16(tbase + (uint32_t)(C0 + 32), va);
        fold(vb, std::integral_constant<int, C0 + 16>{});
      };
      step32(std::integral_constant<int, 0>{});   step32(std::integral_constant<int, 32>{});
      step32(std::integral_constant<int, 64>{});  step32(std::integral_constant<int, 96>{});
      step32(std::integral_constant<int, 128>{}); step32(std::integral_constant<int, 160>{});
      step32(std::integral_constant<int, 192>{}); step32(std::integral_constant<int, 224>{});
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty[acc]));
      const int oy = ty * TH + m / TW, ox = tx * TW + m % TW;
      if (oy < p.H && ox < p.W) {
        float* o = p.out + (((size_t)n * p.H + oy) * p.W + ox) * p.head_C;
#pragma unroll
        for (int c = 0; c < HC; ++c) o[c] = hacc[c] + kc_.head_b[c];
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace

int dec_tail_supported(int C, int Cout, int head_C, int H, int W) {
  return bdrc_tma_encode_fn() != nullptr && Cout == 256 && head_C >= 1 && head_C <= MAX_HEAD && C % 8 == 0 && C <= 512 &&
         H % TH == 0 && W % TW == 0;
}

int dec_tail_launch(bdrc_ctx* ctx, const bf16* x, int N, int H, int W, int C, const float* dw_w9c, const ConvWeights& pw,
                    const float* bias_host, const float* head_w_host, const float* head_b_host, int head_C, float* out,
                    cudaStream_t s) {
  // bias_host [256], head_w_host [head_C][256] (bf16-rounded values as fp32), head_b_host [head_C]: HOST copies, they
  // travel to the kernel as parameters
  BDRC_REQUIRE(ctx, dec_tail_supported(C, pw.Cout, head_C, H, W) && pw.R == 1 && pw.S == 1 && pw.Cin == C && bias_host &&
                        head_w_host && head_b_host,
               "dec_tail: unsupported shape");
  BdrcEncodeTiledFn enc = bdrc_tma_encode_fn();
  const int chunks = pw.Cin_pad / KC;
  CUtensorMap mapX, mapW;
  cuuint32_t es[4] = {1, 1, 1, 1};
  {
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    cuuint32_t box[4] = {KC, HW, HH, 1};
    CUresult r = enc(&mapX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(x), dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return bdrc_fail(ctx, BDRC_ERR_CUDA, "dec_tail: cuTensorMapEncodeTiled(x) failed: %d", (int)r);
  }
  {
    const cuuint64_t K = (cuuint64_t)pw.Cin_pad;
    cuuint64_t dims[2] = {K, (cuuint64_t)pw.Cout_pad};
    cuuint64_t strides[1] = {K * 2};
    cuuint32_t box[2] = {KC, 256};
    CUresult r = enc(&mapW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pw.w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return bdrc_fail(ctx, BDRC_ERR_CUDA, "dec_tail: cuTensorMapEncodeTiled(w) failed: %d", (int)r);
  }
  DecTailParams p;
  p.n_img = N; p.H = H; p.W = W; p.C = C; p.chunks = chunks;
  p.tiles_x = W / TW; p.tiles_y = H / TH; p.total_tiles = N * p.tiles_x * p.tiles_y;
  p.dw_w = dw_w9c; p.out = out; p.head_C = head_C;
  p.idesc = make_idesc_bf16(128, 256);
  p.errflag = ctx->d_errflag;
  const size_t smem = 1024 + 2 * B_BYTES + 2 * A_BYTES + HS * HALO_BYTES + (size_t)(9 * chunks * KC) * 4 + 21 * 8 + 64;
  BDRC_REQUIRE(ctx, smem <= 226 * 1024, "dec_tail: shared memory budget exceeded");
  const int grid = p.total_tiles < ctx->sm_count ? p.total_tiles : ctx->sm_count;
  auto launch = [&](auto hc_tag) -> int {
    constexpr int HC = decltype(hc_tag)::value;
    auto kern = dec_tail_kernel<HC>;
    TailConsts<HC> k;
    memcpy(k.bias, bias_host, sizeof(k.bias));
    memcpy(k.head_b, head_b_host, sizeof(k.head_b));
    memcpy(k.head_w, head_w_host, sizeof(k.head_w));
    BDRC_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    kern<<<grid, DT_THREADS, smem, s>>>(mapX, mapW, p, k);
    BDRC_CHECK_LAUNCH(ctx);
    return 0;
  };
  switch (head_C) {
    case 1: return launch(std::integral_constant<int, 1>{});
    case 2: return launch(std::integral_constant<int, 2>{});
    case 3: return launch(std::integral_constant<int, 3>{});
    case 4: return launch(std::integral_constant<int, 4>{});
    case 5: return launch(std::integral_constant<int, 5>{});
    case 6: return launch(std::integral_constant<int, 6>{});
    case 7: return launch(std::integral_constant<int, 7>{});
    default: return launch(std::integral_constant<int, 8>{});
  }
}
