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
//                allocator, warps 4..11 = epilogue (TMEM -> registers -> bias/residual/ReLU ->
//                bf16 NHWC; lane quarter x column half per warp).  Persistent grid, one CTA per SM, static round-robin tile schedule.
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int BM = 128;          // pixels per tile (UMMA M, cta_group::1)
constexpr int BK = 64;           // channels per K step: 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 384;   // 4 control warps + 8 epilogue warps
constexpr int MAX_STAGES = 8;
constexpr int MAX_G = 4;           // pixel tiles per super tile (interleaved accumulators)
constexpr uint32_t A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr uint32_t OUT_STAGE_BYTES = BM * 128;     // epilogue staging: 128 rows x 64 bf16

struct TcParams {
  int tiles_n, tiles_w, tiles_h, n_img, total_tiles;
  int BW, BH, BN, stages, G;
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
  int out_ld, out_coff, out_fp32, relu, out_tma;
  int* errflag;
  int dbg_skip;   // -DBDRC_EXPERIMENTS builds only -- bit 0: skip A loads, 1: skip B loads, 2: skip epilogue math + I/O,
                  // 3: skip the TMA stores, 4: skip the residual loads; wide kernel -- 5 (32): skip the epilogue,
                  // 6 (64): skip the MMAs, 7 (128): skip the residual, 8 (256): skip the TMA stores
};

using namespace tcptx;

// Timing experiments (skip loads / epilogue, alternative variants) exist only in -DBDRC_EXPERIMENTS builds; the
// product build compiles every one of these branches out.
#ifdef BDRC_EXPERIMENTS
#define DBG(p, bits) ((p).dbg_skip & (bits))
#else
#define DBG(p, bits) 0
#endif

// ConvArgs.pdl launches a conv kernel with programmatic dependent launch (its barrier init / TMEM allocation
// overlaps the tail of the previous kernel).  Measured on B200: the short Easter2 chain gains 3 % at 64 lines
// (69.5 k -> 71.8 k lines/s) and asks for it; the segmentation forward does NOT -- 8 queued C2 steps (1456
// back-to-back launches) drop from 515 to 199 pages/s, with deep queues of cluster launches the early-resident
// CTAs of kernel N+1 hold back kernel N's tail.

__device__ __forceinline__ int floordiv(int a, int b) {
  int q = a / b;
  return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

// ---------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------
// A CTA works on "super tiles": G consecutive 128-pixel tiles x one BN-column tile.  The G pixel
// tiles share one B (weight) tile per K step and own G independent TMEM accumulators; their MMAs
// are issued interleaved (k-step outer, tile inner), so that back-to-back tcgen05.mma instructions
// never accumulate into the same TMEM columns (a dependent accumulate costs ~130 cycles on the
// tensor pipe whatever N is: measured, see profiles/README.md).  G = 4/2/1 for BN <= 64/128/256.
__device__ __forceinline__ void decode_mtile(const TcParams& p, int mt, int& img, int& oy0, int& ox0) {
  const int wt = mt % p.tiles_w;
  mt /= p.tiles_w;
  const int ht = mt % p.tiles_h;
  img = mt / p.tiles_h;
  ox0 = wt * p.BW;
  oy0 = ht * p.BH;
}

// ---------------------------------------------------------------------------------------
// epilogue (warps 4..11), shared by the kernels of this file
// ---------------------------------------------------------------------------------------
// 8 warps in two groups of 4 (warp % 4 = TMEM lane quarter, hardware rule).  The work of one
// super tile is a list of units (pixel tile g, 64-column group j), dealt round-robin to the two
// groups.  bf16 outputs are staged in shared memory in the SWIZZLE_128B layout and written with
// one TMA tensor store per unit (full 128-byte lines; a 16-bytes-per-thread direct store pattern
// cost 2x the MMA time: measured).  The residual of this thread's rows is prefetched into
// registers BEFORE waiting for the accumulators.
// Super tiles st_begin, st_begin + st_step, ... < st_end; `stage_out` = 2 groups x 2 staging buffers.
template <int CTAS, int NBUF>
__device__ __forceinline__ void epilogue_loop(const TcParams& p, const CUtensorMap* tmapO_ptr, uint8_t* stage_out_base,
                                              uint64_t* tfull_bar, uint64_t* tempty_bar, uint32_t tmem_base,
                                              uint32_t cta_rank, int st_begin, int st_end, int st_step) {
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int G = p.G;
  const CUtensorMap& tmapO = *tmapO_ptr;
  {
    // ===================== epilogue =====================
    // 8 warps in two groups of 4 (warp % 4 = TMEM lane quarter, hardware rule).  The work of one
    // super tile is a list of units (pixel tile g, 64-column group j), dealt round-robin to the two
    // groups.  bf16 outputs are staged in shared memory in the SWIZZLE_128B layout and written with
    // one TMA tensor store per unit (full 128-byte lines; a 16-bytes-per-thread direct store pattern
    // cost 2x the MMA time: measured).  The residual of this thread's rows is prefetched into
    // registers BEFORE waiting for the accumulators.
    const int wrank = warp & 3;
    const int egrp = (warp - 4) >> 2;
    const int ngrp64 = (p.BN + 63) / 64;
    const int units = G * ngrp64;              // <= 4 by construction (host)
    const int m = wrank * 32 + lane;           // row of the tile = pixel
    const int hl = m / p.BW, wl = m % p.BW;
    const bool res_vec = p.residual != nullptr && (p.res_ld & 7) == 0;
    const bool issuer = (threadIdx.x == 128 + 128 * egrp);
    uint8_t* stage_out = stage_out_base + (size_t)egrp * NBUF * OUT_STAGE_BYTES;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t use = 0;                          // staging buffer toggle
    const uint32_t tempty0 = CTAS == 2 ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : smem_u32(&tempty_bar[0]);
    for (int st = st_begin; st < st_end; st += st_step) {
      const int nt = st % p.tiles_n;
      const int mg = (st / p.tiles_n) * CTAS + (int)cta_rank;
      const int n0 = nt * p.BN;
      // this group's units: u = egrp, egrp + 2  ->  local index ul = 0, 1
      size_t pix[2];
      bool valid[2];
      int imgs[2], oxs[2], oys[2], cg0[2];
#pragma unroll
      for (int ul = 0; ul < 2; ++ul) {
        const int u = egrp + 2 * ul;
        pix[ul] = 0; valid[ul] = false; imgs[ul] = 0; oxs[ul] = 0; oys[ul] = 0; cg0[ul] = 0;
        if (u < units) {
          const int g = u / ngrp64;
          cg0[ul] = (u - g * ngrp64) * 64;
          int img, oy0, ox0;
          decode_mtile(p, mg * G + g, img, oy0, ox0);
          const int ox = ox0 + wl, oy = oy0 + hl;
          valid[ul] = (ox < p.Wout) && (oy < p.Hout);
          pix[ul] = ((size_t)img * p.Hout + oy) * p.Wout + ox;
          imgs[ul] = img; oxs[ul] = ox0; oys[ul] = oy0;
        }
      }
      uint4 rres[16];
      if (res_vec && !DBG(p, 4 | 16)) {
#pragma unroll
        for (int ul = 0; ul < 2; ++ul) {
          if (egrp + 2 * ul < units && valid[ul]) {
            const bf16* rp = p.residual + pix[ul] * p.res_ld + n0 + cg0[ul];
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (n0 + cg0[ul] + q * 8 + 8 <= p.Cout)
                rres[ul * 8 + q] = *reinterpret_cast<const uint4*>(rp + q * 8);   // plain load: may alias out
          }
        }
      }

      mbar_wait_relaxed(smem_u32(&tfull_bar[acc]), acc_phase, p.errflag, 4);   // sleeps between polls: the other group's epilogue warps need the issue slots
      tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(wrank * 32) << 16) + (uint32_t)(acc * G * p.BN);
#pragma unroll
      for (int ul = 0; ul < 2; ++ul) {
        const int u = egrp + 2 * ul;
        if (u >= units) break;
        const int g = u / ngrp64;
        uint8_t* sbuf = stage_out + (size_t)(use & (NBUF - 1)) * OUT_STAGE_BYTES;
        if (p.out_tma) {
          // the TMA store that last read this staging buffer must have finished reading it
          if (issuer) {
            if (NBUF == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + egrp) : "memory");
        }
        const float* bias = p.bias + (size_t)imgs[ul] * p.bias_img_stride + n0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c0 = cg0[ul] + q * 16;
          if (c0 >= p.BN) break;
          uint32_t v[16];
          tc_ld16(tbase + (uint32_t)(g * p.BN + c0), v);
          tc_wait_ld();
          if (DBG(p, 4)) continue;
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c0 + j));
            f[j + 0] = __uint_as_float(v[j + 0]) + b4.x;
            f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
            f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
            f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
          }
          if (p.residual && valid[ul]) {
#pragma unroll
            for (int hq = 0; hq < 2; ++hq) {
              if (res_vec && n0 + c0 + hq * 8 + 8 <= p.Cout) {
                const uint4 rv = rres[ul * 8 + q * 2 + hq];
                const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&rw[j]);
                  f[hq * 8 + 2 * j] += __low2float(h);
                  f[hq * 8 + 2 * j + 1] += __high2float(h);
                }
              } else {
                const bf16* rp = p.residual + pix[ul] * p.res_ld + n0 + c0 + hq * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j)   // static indices only: keeps f[] in registers
                  if (n0 + c0 + hq * 8 + j < p.Cout) f[hq * 8 + j] += __bfloat162float(rp[j]);
              }
            }
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.0f);
          }
          if (p.out_tma) {
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
              w[j] = *reinterpret_cast<const uint32_t*>(&h);
            }
            // row m, 16-byte chunks 2q and 2q+1 of the 128-byte row, XOR-swizzled with (m & 7)
            uint8_t* rowp = sbuf + (size_t)m * 128;
            *reinterpret_cast<uint4*>(rowp + (((2 * q) ^ (m & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(rowp + (((2 * q + 1) ^ (m & 7)) << 4)) = make_uint4(w[4], w[5], w[6], w[7]);
          } else if (valid[ul] && n0 + c0 < p.Cout) {
            const bool full16 = (n0 + c0 + 16 <= p.Cout);
            if (p.out_fp32) {
              float* op = reinterpret_cast<float*>(p.out) + pix[ul] * p.out_ld + p.out_coff + n0 + c0;
              if (full16 && ((p.out_ld | p.out_coff) & 3) == 0) {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                  *reinterpret_cast<float4*>(op + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (n0 + c0 + j < p.Cout) op[j] = f[j];
              }
            } else {
              bf16* op = reinterpret_cast<bf16*>(p.out) + pix[ul] * p.out_ld + p.out_coff + n0 + c0;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (n0 + c0 + j < p.Cout) op[j] = __float2bfloat16(f[j]);
            }
          }
        }
        if (p.out_tma) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(1 + egrp) : "memory");
          if (issuer && !DBG(p, 4 | 8)) {
            asm volatile(
                "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                ::"l"(reinterpret_cast<uint64_t>(&tmapO)), "r"(smem_u32(sbuf)), "r"(n0 + cg0[ul]), "r"(oxs[ul]),
                "r"(oys[ul]), "r"(imgs[ul])
                : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ++use;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTAS == 2) mbar_arrive_cluster(tempty0 + (uint32_t)acc * 8u); else mbar_arrive(tempty0 + (uint32_t)acc * 8u);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (p.out_tma && issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// CTAS = 1: one CTA per SM works alone (cta_group::1, M = 128).
// CTAS = 2: the two CTAs of a cluster form a pair (cta_group::2, M = 256).  Each CTA loads the A slabs of its
//   own G pixel tiles and HALF of the B tile; the leader (cluster rank 0) issues one MMA for both, which reads
//   A and the B half from each CTA's own shared memory -- 8 KB instead of 12 KB of operand reads per CTA per
//   128-cycle MMA at N = 256, which is what the shared-memory operand path (~64 B/clk/SM, measured: the
//   single-CTA kernel tops out at 67 % tensor-pipe activity with all loads and the epilogue disabled) can feed.
template <int CTAS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
               const __grid_constant__ CUtensorMap tmapO, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment is required by SWIZZLE_128B
  uint8_t* smem = smem_raw + ((1024u - (tcptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int G = p.G;
  const uint32_t cta_rank = CTAS == 2 ? cluster_ctarank() : 0u;
  const int unit_id = CTAS == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // CTA or CTA pair
  const int n_units = CTAS == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const uint32_t b_stage_bytes = (uint32_t)(p.BN / CTAS) * BK * 2;
  const uint32_t a_bytes = (uint32_t)G * A_STAGE_BYTES;
  const uint32_t stage_bytes = a_bytes + b_stage_bytes;
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
      mbar_init(smem_u32(&tempty_bar[i]), 8 * CTAS);  // one arrive per epilogue warp (of both CTAs on the leader)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    if (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                   ::"r"(smem_u32(tmem_slot)), "r"(p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above touched only parameters, shared memory and TMEM; the activations are the previous kernel's
  griddep_launch();
  griddep_wait();

  const int k_iters = p.R * p.S * p.kchunks;

  if (warp == 0) {
    // ===================== TMA producer (whole warp converged, one elected lane issues) ==========
    {
      int stage = 0;
      uint32_t phase = 0;
      // in a pair both CTAs signal the LEADER's full barrier; only the leader arms it (for both CTAs' bytes)
      const uint32_t full0 = CTAS == 2 ? mapa_u32(smem_u32(&full_bar[0]), 0) : smem_u32(&full_bar[0]);
      for (int st = unit_id; st < p.total_tiles; st += n_units) {
        const int nt = st % p.tiles_n;
        const int mg = st / p.tiles_n;
        const int n0 = nt * p.BN + (int)cta_rank * (p.BN / CTAS);
        int img[MAX_G], oy0[MAX_G], ox0[MAX_G];
#pragma unroll
        for (int g = 0; g < MAX_G; ++g)
          if (g < G) decode_mtile(p, (mg * CTAS + (int)cta_rank) * G + g, img[g], oy0[g], ox0[g]);
        for (int r = 0; r < p.R; ++r) {
          const int th = r * p.dil_h - p.pad_h;
          const int par_h = ((th % p.stride_h) + p.stride_h) % p.stride_h;
          const int off_h = floordiv(th - par_h, p.stride_h);
          for (int s = 0; s < p.S; ++s) {
            const int tw = s * p.dil_w - p.pad_w;
            const int par_w = ((tw % p.stride_w) + p.stride_w) % p.stride_w;
            const int off_w = floordiv(tw - par_w, p.stride_w);
            for (int kc = 0; kc < p.kchunks; ++kc) {
              mbar_wait_relaxed(smem_u32(&empty_bar[stage]), phase ^ 1, p.errflag, 1);
              const uint32_t fb = full0 + (uint32_t)stage * 8u;
              const uint32_t a_dst = smem_u32(smem + (size_t)stage * stage_bytes);
              if (CTAS == 2) {
                if (elect_one()) {
                  if (cta_rank == 0) mbar_arrive_expect_tx(fb, 2u * (a_bytes + b_stage_bytes));
#pragma unroll
                  for (int g = 0; g < MAX_G; ++g)
                    if (g < G)
                      tma_load_5d_2sm(a_dst + (uint32_t)g * A_STAGE_BYTES, &tmapA, fb, par_w * p.par_c + kc * BK,
                                      ox0[g] + off_w, par_h, oy0[g] + off_h, img[g]);
                  tma_load_2d_2sm(a_dst + a_bytes, &tmapB, fb, ((r * p.S + s) * p.kchunks + kc) * BK, n0);
                }
              } else {
                const uint32_t bytes = (DBG(p, 1) ? 0u : a_bytes) + (DBG(p, 2) ? 0u : b_stage_bytes);
                if (elect_one()) {
                  if (bytes) mbar_arrive_expect_tx(fb, bytes); else mbar_arrive(fb);
                  if (!DBG(p, 1)) {
#pragma unroll
                    for (int g = 0; g < MAX_G; ++g)
                      if (g < G)
                        tma_load_5d(a_dst + (uint32_t)g * A_STAGE_BYTES, &tmapA, fb, par_w * p.par_c + kc * BK,
                                    ox0[g] + off_w, par_h, oy0[g] + off_h, img[g]);
                  }
                  if (!DBG(p, 2))
                    tma_load_2d(a_dst + a_bytes, &tmapB, fb, ((r * p.S + s) * p.kchunks + kc) * BK, n0);
                }
              }
              __syncwarp();
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp converged, one elected lane issues) ============
    if (cta_rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int st = unit_id; st < p.total_tiles; st += n_units) {
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1, p.errflag, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * G * p.BN);
        for (int it = 0; it < k_iters; ++it) {
          mbar_wait(smem_u32(&full_bar[stage]), phase, p.errflag, 3);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t bdesc = make_sdesc(a_addr + a_bytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 16 bf16 = 32 B along K inside the 128-byte swizzle row: +2 in 16-B units
#pragma unroll
              for (int g = 0; g < MAX_G; ++g)
                if (g < G) {
                  if (CTAS == 2)
                    tc_mma_bf16_2sm(d_tmem + (uint32_t)(g * p.BN),
                                    make_sdesc(a_addr + (uint32_t)g * A_STAGE_BYTES) + (uint64_t)(2 * k),
                                    bdesc + (uint64_t)(2 * k), p.idesc, (it > 0 || k > 0) ? 1u : 0u);
                  else
                    tc_mma_bf16(d_tmem + (uint32_t)(g * p.BN),
                                make_sdesc(a_addr + (uint32_t)g * A_STAGE_BYTES) + (uint64_t)(2 * k),
                                bdesc + (uint64_t)(2 * k), p.idesc, (it > 0 || k > 0) ? 1u : 0u);
                }
            }
            // frees the smem slot (in both CTAs of a pair) when the MMAs retire
            if (CTAS == 2) tc_commit_2sm(smem_u32(&empty_bar[stage])); else tc_commit(smem_u32(&empty_bar[stage]));
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) {   // accumulators complete -> epilogue (of both CTAs)
          if (CTAS == 2) tc_commit_2sm(smem_u32(&tfull_bar[acc])); else tc_commit(smem_u32(&tfull_bar[acc]));
        }
        __syncwarp();
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    epilogue_loop<CTAS, 2>(p, &tmapO, smem + (size_t)p.stages * stage_bytes + 1024, tfull_bar, tempty_bar, tmem_base,
                        cta_rank, unit_id, p.total_tiles, n_units);
  }

  tc_fence_before();
  if (CTAS == 2) cluster_sync_all(); else __syncthreads();   // the peer may still signal barriers in this CTA
  if (warp == 2) {
    tc_fence_after();
    if (CTAS == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// "slab" kernel: 3x3 stride-1 convolutions on 128-pixel-wide maps with Cin = Cout = 64 (ResNet layer 1)
// ---------------------------------------------------------------------------------------
// The generic kernel re-loads the 128 x 64 A slab of a pixel tile for each of the 9 filter taps and, at
// N = 64, spends 72 KB of L2 -> SMEM traffic per 512 MMA clocks (~140 B/clk/SM against the ~50 the L2 can
// deliver): these layers ran at 539 TFLOP/s.  Here an M tile is one full image row, the input rows are
// streamed ONCE through a ring of shared-memory slots (130 pixels x 128 B: left/right halo by TMA
// out-of-bounds fill), and the A operand of tap (r, s) is simply the slot of input row y + r - 1 with its
// start address advanced by s pixels (s x 128 B).  The 128-byte swizzle is a function of the absolute
// shared-memory address bits (measured: the descriptor's base-offset field must stay 0 for such a start; with
// base offset = s the results are wrong), so a window that starts at any 128-byte row reads what TMA wrote.  All 9 weight tiles stay resident (72 KB).  A CTA owns a contiguous range
// of "super rows" (4 output rows = 4 interleaved accumulators); consecutive super rows reuse 2 of their 6
// input rows, slots are released tap-row by tap-row with tcgen05.commit so the ring (7 slots) never stalls.
// One CTA per SM, 7 slots, whole weight tiles resident.  (A CTA-pair variant was built and measured SLOWER, 608
// against 674 TFLOP/s at 148 tiles: at N = 64 the operand path is bound by the 4 KB of A per MMA, which pairing
// does not reduce, and the pair adds a coupled accumulator hand-off.  It is not in the tree any more.)
constexpr uint32_t SLAB_ROW_BYTES = 17 * 1024;        // 130 px x 128 B = 16640, padded to 1024
constexpr uint32_t SLAB_ROW_TX = 130 * 128;
constexpr int SLAB_SLOTS = 7;
constexpr uint32_t SLAB_B_TAP_BYTES = 64 * 128;       // 8 KB per filter tap
constexpr uint32_t SLAB_B_BYTES = 9 * SLAB_B_TAP_BYTES;
constexpr size_t SLAB_SMEM = 1024 + (size_t)SLAB_SLOTS * SLAB_ROW_BYTES + SLAB_B_BYTES + 1024 + 2 * OUT_STAGE_BYTES;

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_slab_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
                    const __grid_constant__ CUtensorMap tmapO, const TcParams p) {
  constexpr int SLOTS = SLAB_SLOTS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tcptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* rows = smem;
  uint8_t* bmat = rows + (size_t)SLOTS * SLAB_ROW_BYTES;
  uint8_t* bar_base = bmat + SLAB_B_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + SLOTS;
  uint64_t* bfull_bar = empty_bar + SLOTS;
  uint64_t* tfull_bar = bfull_bar + 1;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* stage_out = bar_base + 1024;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int unit_id = (int)blockIdx.x;
  const int n_units = (int)gridDim.x;
  const int tiles_per_img = p.Hout / 4;
  // contiguous range of super rows (4 output rows) for this CTA
  const int st_begin = (int)(((long long)p.total_tiles * unit_id) / n_units);
  const int st_end = (int)(((long long)p.total_tiles * (unit_id + 1)) / n_units);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapB)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < SLOTS; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    mbar_init(smem_u32(bfull_bar), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tfull_bar[i]), 1);
      mbar_init(smem_u32(&tempty_bar[i]), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above touched only parameters, shared memory and TMEM; the activations are the previous kernel's
  griddep_launch();
  griddep_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t full0 = smem_u32(&full_bar[0]);
    const uint32_t bfull = smem_u32(bfull_bar);
    if (elect_one()) {
      mbar_arrive_expect_tx(bfull, SLAB_B_BYTES);
      for (int t = 0; t < 9; ++t) tma_load_2d(smem_u32(bmat) + (uint32_t)t * SLAB_B_TAP_BYTES, &tmapB, bfull, t * BK, 0);
    }
    __syncwarp();
    uint32_t L = 0;   // row-load counter: load L goes to slot L % SLOTS, use (L / SLOTS)
    for (int st = st_begin; st < st_end; ++st) {
      const int img = st / tiles_per_img;
      const int y0 = (st - img * tiles_per_img) * 4;
      // rows y0-1, y0 are the previous super row's last two
      const bool carry = (st > st_begin) && (y0 != 0);
      for (int q = carry ? 2 : 0; q < 6; ++q) {
        const uint32_t slot = L % SLOTS, use = L / SLOTS;
        mbar_wait_relaxed(smem_u32(&empty_bar[slot]), (use & 1) ^ 1, p.errflag, 11);
        if (elect_one()) {
          const uint32_t fb = full0 + slot * 8u;
          mbar_arrive_expect_tx(fb, SLAB_ROW_TX);
          tma_load_4d(smem_u32(rows) + slot * SLAB_ROW_BYTES, &tmapA, fb, 0, -1, y0 - 1 + q, img);
        }
        __syncwarp();
        ++L;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    mbar_wait(smem_u32(bfull_bar), 0, p.errflag, 12);
    uint32_t base = 0, waited = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int st = st_begin; st < st_end; ++st) {
      const int y0 = (st % tiles_per_img) * 4;
      const bool next_carry = (st + 1 < st_end) && (y0 + 4 < p.Hout);
      mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1, p.errflag, 2);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 4 * 64);
      for (int r = 0; r < 3; ++r) {
        while (waited <= base + (uint32_t)r + 3u) {
          mbar_wait(smem_u32(&full_bar[waited % SLOTS]), (waited / SLOTS) & 1, p.errflag, 3);
          ++waited;
        }
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const uint64_t bdesc = make_sdesc(smem_u32(bmat) + (uint32_t)(r * 3 + s) * SLAB_B_TAP_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint32_t slot = (base + (uint32_t)(r + g)) % SLOTS;
                const uint32_t a_addr = smem_u32(rows) + slot * SLAB_ROW_BYTES + (uint32_t)s * 128u;
                tc_mma_bf16(d_tmem + (uint32_t)(g * 64), make_sdesc(a_addr) + (uint64_t)(2 * k),
                            bdesc + (uint64_t)(2 * k), p.idesc, (r | s | k) ? 1u : 0u);
              }
            }
          }
          // release the rows whose last use was this tap row
          auto release = [&](uint32_t ln) { tc_commit(smem_u32(&empty_bar[ln % SLOTS])); };
          if (r < 2) {
            release(base + (uint32_t)r);
          } else {
            release(base + 2u);
            release(base + 3u);
            if (!next_carry) {
              release(base + 4u);
              release(base + 5u);
            }
          }
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(smem_u32(&tfull_bar[acc]));
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      base += next_carry ? 4u : 6u;
    }
  } else if (warp >= 4) {
    epilogue_loop<1, 1>(p, &tmapO, stage_out, tfull_bar, tempty_bar, tmem_base, 0u, st_begin, st_end, 1);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// "wide" kernel: the slab kernel's layers (3x3, stride 1, 64 -> 64, 128-pixel-wide maps) with the three
// HORIZONTAL taps moved from the A operand into the N dimension
// ---------------------------------------------------------------------------------------
// At N = 64 every tcgen05.mma reads 4 KB of A for 2 KB of B and the shared-memory operand path, not the tensor
// pipe, sets the pace (the slab kernel: ~97 clocks per 32-clock MMA, 675-755 TFLOP/s).  Here one MMA multiplies an
// UNSHIFTED input row (128 pixels x 16 channels) with the weights of all three horizontal taps of one filter row,
//     D[p][s*64 + co] += in[o + r - 1][p][ci] * W[co][r][s][ci]          (N = 192, K = 3 rows x 64 channels)
// so the A bytes per FLOP drop to a third (12 MMAs of N = 192 per output row instead of 36 of N = 64), and the
// horizontal shift is applied once per output element in the epilogue:
//     out[o][x][co] = D[x-1][co] + D[x][64 + co] + D[x+1][128 + co]     (zero outside the row = the conv padding)
// The x-1 / x+1 terms come from the neighbouring TMEM lanes: warp shuffles inside a lane quarter, a 64-float
// hand-over through shared memory between the four warps of an epilogue group.  The packed weight tiles of the
// taps (r, 0..2) are consecutive 64-row blocks in shared memory, i.e. already one 192-row K-major B operand.
// An M tile is one image row; input rows stream ONCE through a ring of 16 KB slots (no halo columns), a CTA owns a
// contiguous range of output rows, accumulators (192 columns) are double-buffered in TMEM, one epilogue group
// (4 warps) per buffer.
constexpr int WIDE_SLOTS = 5;
constexpr uint32_t WIDE_ROW_BYTES = 128 * 128;                     // 128 px x 64 ch bf16
constexpr uint32_t WIDE_XCH_BYTES = 2 * 4 * 2 * 64 * 4;            // groups x warps x {left, right} x 64 floats
constexpr uint32_t WIDE_ACC_COLS = 192;
constexpr int WIDE_THREADS = 128 + 512;   // 4 control warps + 16 epilogue warps
constexpr size_t WIDE_SMEM =
    1024 + (size_t)WIDE_SLOTS * WIDE_ROW_BYTES + SLAB_B_BYTES + 1024 + 4 * OUT_STAGE_BYTES + WIDE_XCH_BYTES;

// The residual of a ResNet block is fetched by TMA into the very staging buffer the result leaves from (same
// SWIZZLE_128B layout: a thread reads the chunk it is about to overwrite), one output row ahead per epilogue group.
// Per-thread 16-byte global loads of the residual (the shared epilogue above) made the residual launches of these
// layers 1.56x slower than the plain ones (ncu r02: 301 us against 193 us per 148 tiles at +310 MB of reads).
__global__ void __launch_bounds__(WIDE_THREADS, 1)
conv_tc_wide_kernel(const __grid_constant__ CUtensorMap tmapA, const __grid_constant__ CUtensorMap tmapB,
                    const __grid_constant__ CUtensorMap tmapO, const __grid_constant__ CUtensorMap tmapR,
                    const TcParams p) {
  constexpr int SLOTS = WIDE_SLOTS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tcptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* rows = smem;
  uint8_t* bmat = rows + (size_t)SLOTS * WIDE_ROW_BYTES;
  uint8_t* bar_base = bmat + SLAB_B_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + SLOTS;
  uint64_t* bfull_bar = empty_bar + SLOTS;
  uint64_t* tfull_bar = bfull_bar + 1;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* rfull_bar = tempty_bar + 2;      // [group][buffer]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rfull_bar + 4);
  uint8_t* stage_out = bar_base + 1024;      // [group][buffer] x 16 KB
  float* xch = reinterpret_cast<float*>(stage_out + 4 * OUT_STAGE_BYTES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int H = p.Hout;
  // contiguous range of output rows (flattened image-major) for this CTA
  const int f_begin = (int)(((long long)p.total_tiles * blockIdx.x) / gridDim.x);
  const int f_end = (int)(((long long)p.total_tiles * (blockIdx.x + 1)) / gridDim.x);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmapB)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < SLOTS; ++i) {
      mbar_init(smem_u32(&full_bar[i]), 1);
      mbar_init(smem_u32(&empty_bar[i]), 1);
    }
    mbar_init(smem_u32(bfull_bar), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&tfull_bar[i]), 1);
      mbar_init(smem_u32(&tempty_bar[i]), 8);   // the eight warps of the buffer's epilogue group
    }
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&rfull_bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();
  griddep_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    const uint32_t full0 = smem_u32(&full_bar[0]);
    const uint32_t bfull = smem_u32(bfull_bar);
    if (elect_one()) {
      mbar_arrive_expect_tx(bfull, SLAB_B_BYTES);
      for (int t = 0; t < 9; ++t) tma_load_2d(smem_u32(bmat) + (uint32_t)t * SLAB_B_TAP_BYTES, &tmapB, bfull, t * BK, 0);
    }
    __syncwarp();
    uint32_t L = 0;   // row-load counter: load L goes to slot L % SLOTS, use L / SLOTS
    for (int f = f_begin; f < f_end; ++f) {
      const int img = f / H, o = f - img * H;
      const bool new_seg = (f == f_begin) || (o == 0);
      const int y_lo = new_seg ? (o > 0 ? o - 1 : 0) : o + 1;
      const int y_hi = o + 1 < H ? o + 1 : H - 1;
      for (int y = y_lo; y <= y_hi; ++y) {
        const uint32_t slot = L % SLOTS, use = L / SLOTS;
        mbar_wait_relaxed(smem_u32(&empty_bar[slot]), (use & 1) ^ 1, p.errflag, 11);
        if (elect_one()) {
          const uint32_t fb = full0 + slot * 8u;
          mbar_arrive_expect_tx(fb, WIDE_ROW_BYTES);
          tma_load_4d(smem_u32(rows) + slot * WIDE_ROW_BYTES, &tmapA, fb, 0, 0, y, img);
        }
        __syncwarp();
        ++L;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    mbar_wait(smem_u32(bfull_bar), 0, p.errflag, 12);
    uint32_t next_L = 0, seg_base = 0, waited = 0;
    int seg_row0 = 0;
    for (int f = f_begin; f < f_end; ++f) {
      const int img = f / H, o = f - img * H;
      if ((f == f_begin) || (o == 0)) {
        seg_base = next_L;
        seg_row0 = o > 0 ? o - 1 : 0;
        next_L += (o > 0 ? 1u : 0u) + 1u + (o + 1 < H ? 1u : 0u);
      } else {
        next_L += (o + 1 < H ? 1u : 0u);
      }
      const int acc = (f - f_begin) & 1;
      const uint32_t acc_use = (uint32_t)(f - f_begin) >> 1;
      mbar_wait(smem_u32(&tempty_bar[acc]), (acc_use & 1) ^ 1, p.errflag, 2);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)acc * WIDE_ACC_COLS;
      uint32_t accumulate = 0;
      for (int r = 0; r < 3; ++r) {
        const int y = o + r - 1;
        if (y < 0 || y >= H) continue;
        const uint32_t li = seg_base + (uint32_t)(y - seg_row0);
        while (waited <= li) {
          mbar_wait(smem_u32(&full_bar[waited % SLOTS]), (waited / SLOTS) & 1, p.errflag, 3);
          ++waited;
        }
        tc_fence_after();
        if (elect_one()) {
          const uint64_t adesc = make_sdesc(smem_u32(rows) + (li % SLOTS) * WIDE_ROW_BYTES);
          const uint64_t bdesc = make_sdesc(smem_u32(bmat) + (uint32_t)(r * 3) * SLAB_B_TAP_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            if (!DBG(p, 64))
              tc_mma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), p.idesc, (accumulate | k) ? 1u : 0u);
        }
        __syncwarp();
        accumulate = 1;
      }
      if (elect_one()) {
        // rows whose last consumer was this output row go back to the producer when the MMAs retire
        if (o > 0) tc_commit(smem_u32(&empty_bar[(seg_base + (uint32_t)(o - 1 - seg_row0)) % SLOTS]));
        if ((f + 1 == f_end) || (o == H - 1)) {
          tc_commit(smem_u32(&empty_bar[(seg_base + (uint32_t)(o - seg_row0)) % SLOTS]));
          if (o + 1 < H) tc_commit(smem_u32(&empty_bar[(seg_base + (uint32_t)(o + 1 - seg_row0)) % SLOTS]));
        }
        tc_commit(smem_u32(&tfull_bar[acc]));
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===================== epilogue: group e (8 warps) drains accumulator buffer e =====================
    // a warp owns a TMEM lane quarter (warp % 4, hardware rule) and one half of the 64 output channels: four warps per
    // scheduler hide the tcgen05.ld / shuffle / shared-memory latencies of one another (with four warps per buffer the
    // epilogue, 3.7 k clocks per row, was slower than the 1.6 k clocks of MMAs it has to keep up with)
    const int wrank = warp & 3;
    const int egrp = ((warp - 4) >> 2) & 1;
    const int half = (warp - 4) >> 3;
    const int m = wrank * 32 + lane;                 // pixel x of the row = TMEM lane
    const bool issuer = (threadIdx.x == 128 + 128 * egrp);
    const bool has_res = p.residual != nullptr && !DBG(p, 128);
    uint8_t* sbuf0 = stage_out + (size_t)egrp * 2 * OUT_STAGE_BYTES;
    float* xl_mine = xch + ((egrp * 4 + wrank) * 2 + 0) * 64;      // lane 31's s = 0 terms, for the warp to the right
    float* xr_mine = xch + ((egrp * 4 + wrank) * 2 + 1) * 64;      // lane 0's s = 2 terms, for the warp to the left
    const float* xl_nb = xch + ((egrp * 4 + (wrank > 0 ? wrank - 1 : 0)) * 2 + 0) * 64;
    const float* xr_nb = xch + ((egrp * 4 + (wrank < 3 ? wrank + 1 : 3)) * 2 + 1) * 64;
    const uint32_t rfull0 = smem_u32(&rfull_bar[egrp * 2]);
    auto load_residual = [&](int f, uint32_t buf) {   // issuer only
      const int img = f / H, o = f - img * H;
      mbar_arrive_expect_tx(rfull0 + buf * 8u, OUT_STAGE_BYTES);
      tma_load_4d(smem_u32(sbuf0 + (size_t)buf * OUT_STAGE_BYTES), &tmapR, rfull0 + buf * 8u, 0, 0, o, img);
    };
    if (has_res && issuer && f_begin + egrp < f_end) load_residual(f_begin + egrp, 0u);
    uint32_t i = 0;
    for (int f = f_begin + egrp; f < f_end; f += 2, ++i) {
      const uint32_t buf = i & 1u;
      const int img = f / H, o = f - img * H;
      uint8_t* sbuf = sbuf0 + (size_t)buf * OUT_STAGE_BYTES;
      mbar_wait_relaxed(smem_u32(&tfull_bar[egrp]), i & 1u, p.errflag, 4);
      tc_fence_after();
      if (DBG(p, 32)) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[egrp]));
        continue;
      }
      const uint32_t tbase = tmem_base + ((uint32_t)(wrank * 32) << 16) + (uint32_t)egrp * WIDE_ACC_COLS;
      if (issuer) {
        if (has_res) {
          // the store of this group's previous row has left the other buffer: fetch the next row's residual into it
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          if (f + 2 < f_end) load_residual(f + 2, buf ^ 1u);
        } else {
          // this buffer was last read by the store two rows back
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
      }
      asm volatile("bar.sync %0, 256;" ::"r"(1 + egrp) : "memory");
      if (has_res) mbar_wait_relaxed(rfull0 + buf * 8u, (i >> 1) & 1u, p.errflag, 5);
      // ---- shifted sum + bias (+ residual, ReLU) -> bf16 -> swizzled staging, 16 channels at a time ----
      // TMEM reads are the scarce resource here (~100 B/clk/SM, shared with the MMAs' accumulator traffic): every
      // accumulator column is read exactly once, the boundary lanes hand their terms over through shared memory
      // between the load and the arithmetic (a named barrier of the four warps of this channel half).
      const float* bias = p.bias + (size_t)img * p.bias_img_stride;
      uint8_t* rowp = sbuf + (size_t)m * 128;
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const int c = 2 * half + cc;
        uint32_t v0[16], v1[16], v2[16];
        tc_ld16(tbase + (uint32_t)(c * 16), v0);
        tc_ld16(tbase + (uint32_t)(64 + c * 16), v1);
        tc_ld16(tbase + (uint32_t)(128 + c * 16), v2);
        tc_wait_ld();
        if (cc == 1) {
          // accumulator drained (this warp's share): hand the buffer back as early as possible
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[egrp]));
        }
        if (lane == 31) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<uint4*>(xl_mine + c * 16 + j) = make_uint4(v0[j], v0[j + 1], v0[j + 2], v0[j + 3]);
        }
        if (lane == 0) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<uint4*>(xr_mine + c * 16 + j) = make_uint4(v2[j], v2[j + 1], v2[j + 2], v2[j + 3]);
        }
        asm volatile("bar.sync %0, 128;" ::"r"(3 + egrp * 2 + half) : "memory");
        float fo[16];
#pragma unroll
        for (int j4 = 0; j4 < 16; j4 += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c * 16 + j4));
          // the neighbour quarters' hand-over values as unconditional broadcast loads (a per-element load under
          // `if (lane == 0)` serialises 128 divergent shared-memory round trips per row: measured 12 k clocks)
          const float4 l4 = *reinterpret_cast<const float4*>(xl_nb + c * 16 + j4);
          const float4 r4 = *reinterpret_cast<const float4*>(xr_nb + c * 16 + j4);
          const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
          const float ll[4] = {l4.x, l4.y, l4.z, l4.w};
          const float rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = j4 + jj;
            float up = __shfl_up_sync(0xffffffffu, __uint_as_float(v0[j]), 1);
            float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(v2[j]), 1);
            up = lane == 0 ? (wrank > 0 ? ll[jj] : 0.0f) : up;
            dn = lane == 31 ? (wrank < 3 ? rr[jj] : 0.0f) : dn;
            fo[j] = (up + __uint_as_float(v1[j])) + (dn + bb[jj]);
          }
        }
        uint4* c0p = reinterpret_cast<uint4*>(rowp + (((2 * c) ^ (m & 7)) << 4));
        uint4* c1p = reinterpret_cast<uint4*>(rowp + (((2 * c + 1) ^ (m & 7)) << 4));
        if (has_res) {
          const uint4 r0 = *c0p, r1 = *c1p;
          const uint32_t rw[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&rw[j]);
            fo[2 * j] += __low2float(h);
            fo[2 * j + 1] += __high2float(h);
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) fo[j] = fmaxf(fo[j], 0.0f);
        }
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(fo[2 * j], fo[2 * j + 1]);
          w[j] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *c0p = make_uint4(w[0], w[1], w[2], w[3]);
        *c1p = make_uint4(w[4], w[5], w[6], w[7]);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync %0, 256;" ::"r"(1 + egrp) : "memory");
      if (issuer && !DBG(p, 256)) {
        asm volatile(
            "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
            ::"l"(reinterpret_cast<uint64_t>(&tmapO)), "r"(smem_u32(sbuf)), "r"(0), "r"(0), "r"(o), "r"(img)
            : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
typedef BdrcEncodeTiledFn EncodeTiledFn;
EncodeTiledFn get_encode_fn() { return bdrc_tma_encode_fn(); }

struct TcPlan {
  int BW, BH, BN, tiles_w, tiles_h, tiles_n, stages, G, ctas;
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
  // pixel tiles per super tile: as many as TMEM (2 buffers x G x BN <= 512 columns), the epilogue's
  // 4-unit budget (G x ceil(BN/64) <= 4) and divisibility of the pixel-tile count allow
  const long long m_tiles = (long long)pl->tiles_w * pl->tiles_h * a.N;
  int G = 1;
  for (int g = MAX_G; g >= 1; g >>= 1)
    if (2 * g * pl->BN <= 512 && g * ((pl->BN + 63) / 64) <= 4 && m_tiles % g == 0) { G = g; break; }
  pl->G = G;
  // CTA pairs (cta_group::2) whenever the pixel tiles split evenly over the two CTAs
  // and the K loop is long enough to amortise the pair's coupled accumulator hand-off (measured: pairs win
  // on the 3x3 / k-tap layers with N >= 128, lose on 1x1 layers whose tiles are epilogue-bound)
  const int k_iters = w.R * w.S * (w.Cin_pad / BK);
  pl->ctas = (pl->BN >= 128 && pl->BN % 16 == 0 && k_iters >= 16 && m_tiles % (2 * G) == 0) ? 2 : 1;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * G * pl->BN)) cols <<= 1;
  pl->tmem_cols = cols;
  const size_t stage_bytes = (size_t)G * A_STAGE_BYTES + (size_t)(pl->BN / pl->ctas) * BK * 2;
  const size_t out_stage = a.out_fp32 ? 0 : 4 * (size_t)OUT_STAGE_BYTES;   // 2 groups x 2 buffers
  int stages = (int)((224 * 1024 - out_stage) / stage_bytes);
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages < 2) return false;
  pl->stages = stages;
  pl->smem_bytes = 1024 + stages * stage_bytes + 1024 + out_stage;
  return true;
}

}  // namespace

BdrcEncodeTiledFn bdrc_tma_encode_fn() {
  static BdrcEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<BdrcEncodeTiledFn>(p);
  }
  return fn;
}

int conv_tc_supported(const ConvWeights& w, const ConvArgs& a) {
  TcPlan pl;
  return make_plan(w, a, &pl) && get_encode_fn() != nullptr;
}

namespace {

bool slab_eligible(const ConvWeights& w, const ConvArgs& a) {
  return w.R == 3 && w.S == 3 && a.stride_h == 1 && a.stride_w == 1 && a.dil_h == 1 && a.dil_w == 1 && a.pad_h == 1 &&
         a.pad_w == 1 && w.Cin == 64 && w.Cin_pad == 64 && w.Cout_pad == 64 && a.Win == 128 && a.Wout == 128 &&
         a.Hin == a.Hout && a.Hout % 4 == 0 && !a.out_fp32 && a.out_ld % 8 == 0 && a.out_coff % 8 == 0 &&
         a.in_ld % 8 == 0 && a.in_coff % 8 == 0 && a.N >= 1;
}

// Everything one launch needs.  Built once per distinct (weights, ConvArgs) and kept in the context
// (TcCache): a forward re-uses the same few dozen entries call after call, so the steady state encodes no
// tensor maps and derives no tiling on the host.
enum { TC_GENERIC1 = 0, TC_GENERIC2 = 1, TC_SLAB = 2, TC_WIDE = 3 };
struct TcLaunch {
  CUtensorMap mapA, mapB, mapO, mapR;
  TcParams p;
  int kind = TC_GENERIC1;
  unsigned grid = 0;
  size_t smem = 0;
};

// POD key: every field of the problem that reaches a tensor map or TcParams (padding bytes zeroed by memset)
struct TcKey {
  const void *w, *b, *in, *out, *residual, *bias_override;
  int Cout, Cout_pad, Cin, Cin_pad, R, S;
  int N, Hin, Win, in_ld, in_coff, stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, Hout, Wout;
  int bias_img_stride, res_ld, out_ld, out_coff, out_fp32, relu, sm_count, variant;
};

}  // namespace

struct TcCache {
  std::mutex mu;
  std::unordered_map<std::string, TcLaunch> map;
};

void conv_tc_cache_free(bdrc_ctx* ctx) {
  delete reinterpret_cast<TcCache*>(ctx->tc_cache);
  ctx->tc_cache = nullptr;
}

namespace {

int encode(bdrc_ctx* ctx, EncodeTiledFn enc, CUtensorMap* m, int rank, void* base, const cuuint64_t* dims,
           const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapL2promotion l2, const char* what) {
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return bdrc_fail(ctx, BDRC_ERR_CUDA, "cuTensorMapEncodeTiled(%s) failed: %d", what, (int)r);
  return 0;
}

int encode_weights(bdrc_ctx* ctx, EncodeTiledFn enc, CUtensorMap* m, const ConvWeights& w, int box_n) {
  const cuuint64_t K = (cuuint64_t)w.R * w.S * w.Cin_pad;
  cuuint64_t dims[2] = {K, (cuuint64_t)w.Cout_pad};
  cuuint64_t strides[1] = {K * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_n};
  return encode(ctx, enc, m, 2, w.w, dims, strides, box, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "weights");
}

int encode_out(bdrc_ctx* ctx, EncodeTiledFn enc, CUtensorMap* m, const ConvWeights& w, const ConvArgs& a, int BW, int BH) {
  const cuuint64_t ld = (cuuint64_t)a.out_ld;
  cuuint64_t dims[4] = {(cuuint64_t)w.Cout, (cuuint64_t)a.Wout, (cuuint64_t)a.Hout, (cuuint64_t)a.N};
  cuuint64_t strides[3] = {ld * 2, (cuuint64_t)a.Wout * ld * 2, (cuuint64_t)a.Hout * a.Wout * ld * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)BW, (cuuint32_t)BH, 1};
  return encode(ctx, enc, m, 4, reinterpret_cast<bf16*>(a.out) + a.out_coff, dims, strides, box,
                CU_TENSOR_MAP_L2_PROMOTION_NONE, "out");
}

void common_params(bdrc_ctx* ctx, const ConvWeights& w, const ConvArgs& a, TcParams* p) {
  p->n_img = a.N;
  p->par_c = a.in_ld;
  p->Hout = a.Hout; p->Wout = a.Wout; p->Cout = w.Cout;
  p->bias = a.bias_override ? a.bias_override : w.b;
  p->bias_img_stride = a.bias_override ? a.bias_img_stride : 0;
  p->residual = a.residual; p->res_ld = a.res_ld;
  p->out = a.out; p->out_ld = a.out_ld; p->out_coff = a.out_coff; p->out_fp32 = a.out_fp32;
  p->relu = a.relu;
  p->errflag = ctx->d_errflag;
  p->dbg_skip = 0;
#ifdef BDRC_EXPERIMENTS
  if (const char* e = getenv("BDRC_DBG_SKIP")) p->dbg_skip = atoi(e);
#endif
}

int build_slab(bdrc_ctx* ctx, EncodeTiledFn enc, const ConvWeights& w, const ConvArgs& a, TcLaunch* L) {
  {
    const cuuint64_t ld = (cuuint64_t)a.in_ld;
    cuuint64_t dims[4] = {(cuuint64_t)w.Cin, (cuuint64_t)a.Win, (cuuint64_t)a.Hin, (cuuint64_t)a.N};
    cuuint64_t strides[3] = {ld * 2, (cuuint64_t)a.Win * ld * 2, (cuuint64_t)a.Hin * a.Win * ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, 130, 1, 1};
    BDRC_TRY(encode(ctx, enc, &L->mapA, 4, const_cast<bf16*>(a.in) + a.in_coff, dims, strides, box,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "slab A"));
  }
  BDRC_TRY(encode_weights(ctx, enc, &L->mapB, w, 64));
  BDRC_TRY(encode_out(ctx, enc, &L->mapO, w, a, 128, 1));
  TcParams& p = L->p;
  common_params(ctx, w, a, &p);
  p.tiles_n = 1; p.tiles_w = 1; p.tiles_h = a.Hout;
  p.total_tiles = a.N * (a.Hout / 4);
  p.BW = 128; p.BH = 1; p.BN = 64; p.stages = 0; p.G = 4;
  p.R = 3; p.S = 3; p.kchunks = 1;
  p.stride_h = p.stride_w = 1; p.pad_h = p.pad_w = 1; p.dil_h = p.dil_w = 1;
  p.idesc = make_idesc_bf16(BM, 64);
  p.tmem_cols = 512;
  p.out_fp32 = 0;
  p.out_tma = 1;
  L->kind = TC_SLAB;
  L->grid = (unsigned)(p.total_tiles < ctx->sm_count ? p.total_tiles : ctx->sm_count);
  L->smem = SLAB_SMEM;
  return 0;
}

// the slab kernel's layers on the wide kernel: additionally the residual (if any) must allow 16-byte loads
bool wide_eligible(const ConvWeights& w, const ConvArgs& a) {
  return slab_eligible(w, a) && w.Cout == 64 && (a.residual == nullptr || a.res_ld % 8 == 0);
}

int build_wide(bdrc_ctx* ctx, EncodeTiledFn enc, const ConvWeights& w, const ConvArgs& a, TcLaunch* L) {
  {
    const cuuint64_t ld = (cuuint64_t)a.in_ld;
    cuuint64_t dims[4] = {(cuuint64_t)w.Cin, (cuuint64_t)a.Win, (cuuint64_t)a.Hin, (cuuint64_t)a.N};
    cuuint64_t strides[3] = {ld * 2, (cuuint64_t)a.Win * ld * 2, (cuuint64_t)a.Hin * a.Win * ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, 128, 1, 1};
    BDRC_TRY(encode(ctx, enc, &L->mapA, 4, const_cast<bf16*>(a.in) + a.in_coff, dims, strides, box,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "wide A"));
  }
  BDRC_TRY(encode_weights(ctx, enc, &L->mapB, w, 64));
  BDRC_TRY(encode_out(ctx, enc, &L->mapO, w, a, 128, 1));
  L->mapR = L->mapO;
  if (a.residual) {
    const cuuint64_t ld = (cuuint64_t)a.res_ld;
    cuuint64_t dims[4] = {64, (cuuint64_t)a.Wout, (cuuint64_t)a.Hout, (cuuint64_t)a.N};
    cuuint64_t strides[3] = {ld * 2, (cuuint64_t)a.Wout * ld * 2, (cuuint64_t)a.Hout * a.Wout * ld * 2};
    cuuint32_t box[4] = {64, 128, 1, 1};
    BDRC_TRY(encode(ctx, enc, &L->mapR, 4, const_cast<bf16*>(a.residual), dims, strides, box,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "wide residual"));
  }
  TcParams& p = L->p;
  common_params(ctx, w, a, &p);
  p.tiles_n = 1; p.tiles_w = 1; p.tiles_h = a.Hout;
  p.total_tiles = a.N * a.Hout;                      // output rows
  p.BW = 128; p.BH = 1; p.BN = 192; p.stages = 0; p.G = 1;
  p.R = 3; p.S = 3; p.kchunks = 1;
  p.stride_h = p.stride_w = 1; p.pad_h = p.pad_w = 1; p.dil_h = p.dil_w = 1;
  p.idesc = make_idesc_bf16(BM, 192);
  p.tmem_cols = 512;
  p.out_fp32 = 0;
  p.out_tma = 1;
  L->kind = TC_WIDE;
  L->grid = (unsigned)(p.total_tiles < ctx->sm_count ? p.total_tiles : ctx->sm_count);
  L->smem = WIDE_SMEM;
  return 0;
}

int build_generic(bdrc_ctx* ctx, EncodeTiledFn enc, const ConvWeights& w, const ConvArgs& a, TcLaunch* L) {
  TcPlan pl;
  if (!make_plan(w, a, &pl)) return bdrc_fail(ctx, BDRC_ERR_UNSUPPORTED, "conv_tc: unsupported shape");
  const int sw = a.stride_w, sh = a.stride_h;
  {
    const cuuint64_t ld = (cuuint64_t)a.in_ld;
    cuuint64_t dims[5] = {sw == 2 ? 2 * ld : (cuuint64_t)w.Cin, (cuuint64_t)(a.Win / sw),
                          (cuuint64_t)sh, (cuuint64_t)(a.Hin / sh), (cuuint64_t)a.N};
    cuuint64_t strides[4] = {(cuuint64_t)sw * ld * 2, (cuuint64_t)a.Win * ld * 2,
                             (cuuint64_t)sh * a.Win * ld * 2, (cuuint64_t)a.Hin * a.Win * ld * 2};
    cuuint32_t box[5] = {(cuuint32_t)BK, (cuuint32_t)pl.BW, 1, (cuuint32_t)pl.BH, 1};
    BDRC_TRY(encode(ctx, enc, &L->mapA, 5, const_cast<bf16*>(a.in) + a.in_coff, dims, strides, box,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "A"));
  }
  BDRC_TRY(encode_weights(ctx, enc, &L->mapB, w, pl.BN / pl.ctas));
  const int out_tma = (!a.out_fp32 && a.out_ld % 8 == 0 && a.out_coff % 8 == 0) ? 1 : 0;
  L->mapO = L->mapB;
  if (out_tma) BDRC_TRY(encode_out(ctx, enc, &L->mapO, w, a, pl.BW, pl.BH));

  TcParams& p = L->p;
  common_params(ctx, w, a, &p);
  p.tiles_n = pl.tiles_n; p.tiles_w = pl.tiles_w; p.tiles_h = pl.tiles_h;
  p.total_tiles = pl.tiles_n * ((pl.tiles_w * pl.tiles_h * a.N) / (pl.G * pl.ctas));   // super tiles (per CTA / pair)
  p.BW = pl.BW; p.BH = pl.BH; p.BN = pl.BN; p.stages = pl.stages; p.G = pl.G;
  p.R = w.R; p.S = w.S; p.kchunks = w.Cin_pad / BK;
  p.stride_h = sh; p.stride_w = sw; p.pad_h = a.pad_h; p.pad_w = a.pad_w;
  p.dil_h = a.dil_h; p.dil_w = a.dil_w;
  // instruction descriptor: D=f32, A=B=bf16, both K-major, N = BN, M = 128 (256 for a CTA pair)
  p.idesc = make_idesc_bf16(BM * pl.ctas, pl.BN);
  p.tmem_cols = pl.tmem_cols;
  p.out_tma = out_tma;
  L->smem = pl.smem_bytes;
  if (pl.ctas == 2) {
    const int pairs = ctx->sm_count / 2;
    L->kind = TC_GENERIC2;
    L->grid = 2u * (unsigned)(p.total_tiles < pairs ? p.total_tiles : pairs);
  } else {
    L->kind = TC_GENERIC1;
    L->grid = (unsigned)(p.total_tiles < ctx->sm_count ? p.total_tiles : ctx->sm_count);
  }
  return 0;
}

}  // namespace

int conv_tc_launch(bdrc_ctx* ctx, const ConvWeights& w, const ConvArgs& a, cudaStream_t s) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return bdrc_fail(ctx, BDRC_ERR_CUDA, "cuTensorMapEncodeTiled not available");
  static std::once_flag attr_once;
  static cudaError_t attr_rc = cudaSuccess;
  std::call_once(attr_once, [] {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SLAB_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_tc_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WIDE_SMEM);
    attr_rc = e;
  });
  if (attr_rc != cudaSuccess)
    return bdrc_fail(ctx, BDRC_ERR_CUDA, "cudaFuncSetAttribute(conv_tc) -> %s", cudaGetErrorString(attr_rc));

  TcKey key;
  memset(&key, 0, sizeof(key));
  key.w = w.w; key.b = w.b; key.in = a.in; key.out = a.out; key.residual = a.residual; key.bias_override = a.bias_override;
  key.Cout = w.Cout; key.Cout_pad = w.Cout_pad; key.Cin = w.Cin; key.Cin_pad = w.Cin_pad; key.R = w.R; key.S = w.S;
  key.N = a.N; key.Hin = a.Hin; key.Win = a.Win; key.in_ld = a.in_ld; key.in_coff = a.in_coff;
  key.stride_h = a.stride_h; key.stride_w = a.stride_w; key.pad_h = a.pad_h; key.pad_w = a.pad_w;
  key.dil_h = a.dil_h; key.dil_w = a.dil_w; key.Hout = a.Hout; key.Wout = a.Wout;
  key.bias_img_stride = a.bias_img_stride; key.res_ld = a.res_ld; key.out_ld = a.out_ld; key.out_coff = a.out_coff;
  key.out_fp32 = a.out_fp32; key.relu = a.relu; key.sm_count = ctx->sm_count; key.variant = ctx->no_wide_conv ? 1 : 0;
  const std::string k(reinterpret_cast<const char*>(&key), sizeof(key));

  TcLaunch L;
  {
    if (!ctx->tc_cache) {
      static std::mutex create_mu;
      std::lock_guard<std::mutex> g(create_mu);
      if (!ctx->tc_cache) ctx->tc_cache = new TcCache();
    }
    TcCache* cache = reinterpret_cast<TcCache*>(ctx->tc_cache);
    std::lock_guard<std::mutex> g(cache->mu);
    auto it = cache->map.find(k);
    if (it == cache->map.end()) {
      TcLaunch fresh;
      if (!ctx->no_wide_conv && wide_eligible(w, a)) BDRC_TRY(build_wide(ctx, enc, w, a, &fresh));
      else BDRC_TRY(slab_eligible(w, a) ? build_slab(ctx, enc, w, a, &fresh) : build_generic(ctx, enc, w, a, &fresh));
      if (cache->map.size() >= 8192) cache->map.clear();   // callers that never repeat a buffer address
      it = cache->map.emplace(k, fresh).first;
      ctx->tc_plans_built++;
    }
    L = it->second;
  }

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(L.grid);
  cfg.blockDim = dim3(L.kind == TC_WIDE ? WIDE_THREADS : NUM_THREADS);
  cfg.dynamicSmemBytes = L.smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (L.kind == TC_GENERIC2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (a.pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (L.kind == TC_GENERIC2)
    BDRC_CUDA(ctx, cudaLaunchKernelEx(&cfg, conv_tc_kernel<2>, L.mapA, L.mapB, L.mapO, L.p));
  else if (L.kind == TC_GENERIC1)
    BDRC_CUDA(ctx, cudaLaunchKernelEx(&cfg, conv_tc_kernel<1>, L.mapA, L.mapB, L.mapO, L.p));
  else if (L.kind == TC_WIDE)
    BDRC_CUDA(ctx, cudaLaunchKernelEx(&cfg, conv_tc_wide_kernel, L.mapA, L.mapB, L.mapO, L.mapR, L.p));
  else
    BDRC_CUDA(ctx, cudaLaunchKernelEx(&cfg, conv_tc_slab_kernel, L.mapA, L.mapB, L.mapO, L.p));
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}
