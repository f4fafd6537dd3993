// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM), operands staged by TMA.
//
// Replaces the onnxruntime Conv nodes of both networks (BDRC/inference.py:152-160, 359-370;
// node list PyTorch-LineInference.ipynb:L233-L469).  One kernel covers conv2d 1x1/3x3 with
// stride 1|2 and any dilation, and conv1d (H = 1) with k = 1..11 and dilation.
//
//   GEMM view :  D[M = N*Hout*Wout pixels][Cout] = sum over (r,s,cin) A[pixel shifted by tap][cin] * W[cout][r][s][cin]
//   A operand :  NHWC bf16 activations.  A 128-pixel x 64-channel slab for one filter tap is ONE
//                TMA box of a 5-D tensor map (c, w, h-parity, h, n); padding halo and ragged edges
//                come back as zeros from TMA out-of-bounds fill, so there is no im2col buffer.
//                Stride-2 layers use the same map with the w/h parity folded into the channel /
//                an extra dimension (pixel pairs are contiguous in NHWC).
//   B operand :  packed weights [Cout_pad][R*S*Cin_pad] bf16, K-major, 2-D TMA box (64 x BN).
//   D         :  fp32 accumulators in TMEM, double-buffered (2 x BN columns) so the epilogue of
//                tile i overlaps the MMAs of tile i+1.
//   Roles     :  warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warp 2 = TMEM
//                allocator, warps 4..7 = epilogue (TMEM -> registers -> bias/residual/ReLU ->
//                bf16 NHWC).  Persistent grid, one CTA per SM, static round-robin tile schedule.
#include "common.cuh"

namespace {

constexpr int BM = 128;          // pixels per tile (UMMA M, cta_group::1)
constexpr int BK = 64;           // channels per K step: 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 256;
constexpr int MAX_STAGES = 8;
constexpr uint32_t A_STAGE_BYTES = BM * BK * 2;  // 16 KB

struct TcParams {
  int tiles_n, tiles_w, tiles_h, n_img, total_tiles;
  int BW, BH, BN, stages;
  int R, S, kchunks;
  int stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
  int par_c;                       // channel offset of the odd-w pixel in the folded view
  int Hout, Wout, Cout;
  uint32_t idesc;
  uint32_t tmem_cols;
  const float* bias;
  int bias_img_stride;
  const bf16* residual;
  int res_ld;
  void* out;
  int out_ld, out_coff, out_fp32, relu;
  int* errflag;
};

// ---------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Bounded wait: a protocol bug must trap, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* errflag, int code) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0x3ff) == 0 && clock64() - t0 > 4000000000LL) {
      if (errflag) atomicExch(errflag, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"):
//   start address >> 4 | LBO = 1 (unused for swizzled K-major) | SBO = 1024 B (8 rows x 128 B)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ int floordiv(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// ---------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
               const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  const uint32_t b_stage_bytes = (uint32_t)p.BN * BK * 2;
  const uint32_t stage_bytes = A_STAGE_BYTES + b_stage_bytes;
  uint8_t* bar_base = smem + (size_t)p.stages * stage_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tfull_bar = empty_bar + MAX_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapB)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tfull_bar[i]), 1);
      mbar_init(smem_u32(&tempty_bar[i]), 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int k_iters = p.R * p.S * p.kchunks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        int t = tile;
        const int nt = t % p.tiles_n; t /= p.tiles_n;
        const int wt = t % p.tiles_w; t /= p.tiles_w;
        const int ht = t % p.tiles_h;
        const int img = t / p.tiles_h;
        const int ox0 = wt * p.BW, oy0 = ht * p.BH, n0 = nt * p.BN;
        for (int r = 0; r < p.R; ++r) {
          const int th = r * p.dil_h - p.pad_h;
          const int par_h = ((th % p.stride_h) + p.stride_h) % p.stride_h;
          const int off_h = floordiv(th - par_h, p.stride_h);
          for (int s = 0; s < p.S; ++s) {
            const int tw = s * p.dil_w - p.pad_w;
            const int par_w = ((tw % p.stride_w) + p.stride_w) % p.stride_w;
            const int off_w = floordiv(tw - par_w, p.stride_w);
            for (int kc = 0; kc < p.kchunks; ++kc) {
              mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1, p.errflag, 1);
              const uint32_t fb = smem_u32(&full_bar[stage]);
              mbar_arrive_expect_tx(fb, stage_bytes);
              const uint32_t a_dst = smem_u32(smem + (size_t)stage * stage_bytes);
              tma_load_5d(a_dst, &tmapA, fb, par_w * p.par_c + kc * BK, ox0 + off_w, par_h,
                          oy0 + off_h, img);
              tma_load_2d(a_dst + A_STAGE_BYTES, &tmapB, fb,
                          ((r * p.S + s) * p.kchunks + kc) * BK, n0);
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1, p.errflag, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.BN);
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(smem_u32(&full_bar[stage]), phase, p.errflag, 3);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = make_sdesc(a_addr);
          const uint64_t bdesc = make_sdesc(a_addr + A_STAGE_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128-byte swizzle row: +2 in 16-B units
            tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc,
                        (it > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit(smem_u32(&empty_bar[stage]));   // frees the smem slot when the MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        tc_commit(smem_u32(&tfull_bar[acc]));       // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int wrank = warp & 3;                // TMEM lane quarter this warp may access
    const int m = wrank * 32 + lane;           // row of the tile = pixel
    const int hl = m / p.BW, wl = m % p.BW;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      int t = tile;
      const int nt = t % p.tiles_n; t /= p.tiles_n;
      const int wt = t % p.tiles_w; t /= p.tiles_w;
      const int ht = t % p.tiles_h;
      const int img = t / p.tiles_h;
      const int ox = wt * p.BW + wl, oy = ht * p.BH + hl, n0 = nt * p.BN;
      const bool valid = (ox < p.Wout) && (oy < p.Hout);
      const size_t pix = ((size_t)img * p.Hout + oy) * p.Wout + ox;
      const float* bias = p.bias + (size_t)img * p.bias_img_stride + n0;

      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase, p.errflag, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(wrank * 32) << 16) + (uint32_t)(acc * p.BN);
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t v[16];
        tc_ld16(taddr + (uint32_t)c0, v);
        tc_wait_ld();
        if (valid && n0 + c0 < p.Cout) {
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c0 + j));
            f[j + 0] = __uint_as_float(v[j + 0]) + b4.x;
            f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
            f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
            f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
          }
          const bool full16 = (n0 + c0 + 16 <= p.Cout);
          if (p.residual) {
            const bf16* rp = p.residual + pix * p.res_ld + n0 + c0;
            if (full16 && (p.res_ld & 7) == 0) {
              // plain loads: the residual may alias the output (in-place shortcut add)
              const uint4 r0 = *reinterpret_cast<const uint4*>(rp);
              const uint4 r1 = *reinterpret_cast<const uint4*>(rp + 8);
              const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&rw[j]);
                f[2 * j] += __low2float(h);
                f[2 * j + 1] += __high2float(h);
              }
            } else {
              for (int j = 0; j < 16 && n0 + c0 + j < p.Cout; ++j) f[j] += __bfloat162float(rp[j]);
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.0f);
          }
          if (p.out_fp32) {
            float* op = reinterpret_cast<float*>(p.out) + pix * p.out_ld + p.out_coff + n0 + c0;
            if (full16 && ((p.out_ld | p.out_coff) & 3) == 0) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(op + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
            } else {
              for (int j = 0; j < 16 && n0 + c0 + j < p.Cout; ++j) op[j] = f[j];
            }
          } else {
            bf16* op = reinterpret_cast<bf16*>(p.out) + pix * p.out_ld + p.out_coff + n0 + c0;
            if (full16 && ((p.out_ld | p.out_coff) & 7) == 0) {
              uint32_t w[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                w[j] = *reinterpret_cast<const uint32_t*>(&h);
              }
              *reinterpret_cast<uint4*>(op) = make_uint4(w[0], w[1], w[2], w[3]);
              *reinterpret_cast<uint4*>(op + 8) = make_uint4(w[4], w[5], w[6], w[7]);
            } else {
              for (int j = 0; j < 16 && n0 + c0 + j < p.Cout; ++j) op[j] = __float2bfloat16(f[j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(p.tmem_cols)
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct TcPlan {
  int BW, BH, BN, tiles_w, tiles_h, tiles_n, stages;
  uint32_t tmem_cols;
  size_t smem_bytes;
};

bool make_plan(const ConvWeights& w, const ConvArgs& a, TcPlan* pl) {
  if (a.in_ld % 8 || a.in_coff % 8) return false;
  if (a.stride_h < 1 || a.stride_h > 2 || a.stride_w < 1 || a.stride_w > 2) return false;
  if (a.stride_w == 2 && (a.in_coff != 0 || a.in_ld != w.Cin || w.Cin % 64 || a.Win % 2))
    return false;
  if (a.stride_h == 2 && (a.Hin % 2)) return false;
  if (w.Cout_pad % 16) return false;
  if (a.Wout >= 128) {
    pl->BW = 128;
    pl->BH = 1;
  } else if (a.Wout == 64 || a.Wout == 32 || a.Wout == 16 || a.Wout == 8) {
    pl->BW = a.Wout;
    pl->BH = 128 / a.Wout;
  } else {
    return false;
  }
  pl->tiles_w = ceil_div(a.Wout, pl->BW);
  pl->tiles_h = ceil_div(a.Hout, pl->BH);
  if (w.Cout_pad <= 256) {
    pl->BN = w.Cout_pad;
  } else if (w.Cout_pad % 256 == 0) {
    pl->BN = 256;
  } else {
    pl->BN = 128;
  }
  pl->tiles_n = ceil_div(w.Cout_pad, pl->BN);
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * pl->BN)) cols <<= 1;
  pl->tmem_cols = cols;
  const size_t stage_bytes = A_STAGE_BYTES + (size_t)pl->BN * BK * 2;
  int stages = (int)((200 * 1024) / stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return false;
  pl->stages = stages;
  pl->smem_bytes = 1024 + stages * stage_bytes + 256;
  return true;
}

}  // namespace

int conv_tc_supported(const ConvWeights& w, const ConvArgs& a) {
  TcPlan pl;
  return make_plan(w, a, &pl) && get_encode_fn() != nullptr;
}

int conv_tc_launch(bdrc_ctx* ctx, const ConvWeights& w, const ConvArgs& a, cudaStream_t s) {
  TcPlan pl;
  if (!make_plan(w, a, &pl)) return bdrc_fail(ctx, BDRC_ERR_UNSUPPORTED, "conv_tc: unsupported shape");
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return bdrc_fail(ctx, BDRC_ERR_CUDA, "cuTensorMapEncodeTiled not available");

  const int sw = a.stride_w, sh = a.stride_h;
  CUtensorMap mapA, mapB;
  {
    const cuuint64_t ld = (cuuint64_t)a.in_ld;
    cuuint64_t dims[5] = {sw == 2 ? 2 * ld : (cuuint64_t)w.Cin, (cuuint64_t)(a.Win / sw),
                          (cuuint64_t)sh, (cuuint64_t)(a.Hin / sh), (cuuint64_t)a.N};
    cuuint64_t strides[4] = {(cuuint64_t)sw * ld * 2, (cuuint64_t)a.Win * ld * 2,
                             (cuuint64_t)sh * a.Win * ld * 2, (cuuint64_t)a.Hin * a.Win * ld * 2};
    cuuint32_t box[5] = {(cuuint32_t)BK, (cuuint32_t)pl.BW, 1, (cuuint32_t)pl.BH, 1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    void* base = const_cast<bf16*>(a.in) + a.in_coff;
    CUresult r = enc(&mapA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return bdrc_fail(ctx, BDRC_ERR_CUDA, "cuTensorMapEncodeTiled(A) failed: %d", (int)r);
  }
  {
    const cuuint64_t K = (cuuint64_t)w.R * w.S * w.Cin_pad;
    cuuint64_t dims[2] = {K, (cuuint64_t)w.Cout_pad};
    cuuint64_t strides[1] = {K * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)pl.BN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w.w, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
      return bdrc_fail(ctx, BDRC_ERR_CUDA, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
  }

  TcParams p;
  p.tiles_n = pl.tiles_n; p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h; p.n_img = a.N;
  p.total_tiles = pl.tiles_n * pl.tiles_w * pl.tiles_h * a.N;
  p.BW = pl.BW; p.BH = pl.BH; p.BN = pl.BN; p.stages = pl.stages;
  p.R = w.R; p.S = w.S; p.kchunks = w.Cin_pad / BK;
  p.stride_h = sh; p.stride_w = sw; p.pad_h = a.pad_h; p.pad_w = a.pad_w;
  p.dil_h = a.dil_h; p.dil_w = a.dil_w;
  p.par_c = a.in_ld;
  p.Hout = a.Hout; p.Wout = a.Wout; p.Cout = w.Cout;
  // instruction descriptor: D=f32, A=B=bf16, both K-major, N = BN, M = 128
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(pl.BN >> 3) << 17) |
            ((uint32_t)(BM >> 4) << 24);
  p.tmem_cols = pl.tmem_cols;
  p.bias = a.bias_override ? a.bias_override : w.b;
  p.bias_img_stride = a.bias_override ? a.bias_img_stride : 0;
  p.residual = a.residual; p.res_ld = a.res_ld;
  p.out = a.out; p.out_ld = a.out_ld; p.out_coff = a.out_coff; p.out_fp32 = a.out_fp32;
  p.relu = a.relu;
  p.errflag = ctx->d_errflag;

  static bool attr_set = false;
  if (!attr_set) {
    BDRC_CUDA(ctx, cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        220 * 1024));
    attr_set = true;
  }
  int grid = p.total_tiles < ctx->sm_count ? p.total_tiles : ctx->sm_count;
  conv_tc_kernel<<<grid, NUM_THREADS, pl.smem_bytes, s>>>(mapA, mapB, p);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}
