// Shared internals of libbdrc_b200.so (not part of the C ABI; see include/bdrc_b200.h).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bdrc_b200.h"

typedef __nv_bfloat16 bf16;

struct bdrc_ctx {
  int device = 0;
  int sm_count = 148;
  char err[512] = {0};
  // per-layer profiling (optional): event pairs recorded around conv launches
  bool profile = false;
  std::vector<cudaEvent_t> prof_ev;        // 2 events per recorded launch
  std::vector<std::string> prof_name;
  std::vector<double> prof_flops;
  size_t prof_used = 0;
  uint64_t launches = 0;                   // kernels launched through this ctx
  int conv_engine = 0;                     // 0 = tcgen05 implicit GEMM, 1 = SIMT (debug cross-check, explicit)
  bool no_dec_tail = false;                // BDRC_NO_DEC_TAIL=1: run the decoder tail as three kernels (cross-check)
  bool no_wide_conv = false;               // BDRC_NO_WIDE_CONV=1: 64-channel 3x3 layers on the slab kernel (cross-check)
  bool allow_simt = false;                 // BDRC_ALLOW_SIMT=1: shapes the tcgen05 engine cannot tile may run SIMT
  uint64_t simt_launches = 0;              // convolutions that ran on the SIMT engine (bdrc_simt_launch_count)
  uint64_t tc_plans_built = 0;             // tensor-map / tiling plans built (cache misses of conv_tc_launch)
  void* tc_cache = nullptr;                // TcCache* (conv_tc.cu)
  int* d_errflag = nullptr;                // device-side watchdog flag
};

int bdrc_fail(bdrc_ctx* ctx, int code, const char* fmt, ...);

#define BDRC_CUDA(ctx, call)                                                        \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess)                                                         \
      return bdrc_fail((ctx), BDRC_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__,  \
                       #call, cudaGetErrorString(e__));                             \
  } while (0)

#define BDRC_CHECK_LAUNCH(ctx)                                                      \
  do {                                                                              \
    (ctx)->launches++;                                                              \
    cudaError_t e__ = cudaGetLastError();                                           \
    if (e__ != cudaSuccess)                                                         \
      return bdrc_fail((ctx), BDRC_ERR_CUDA, "%s:%d launch -> %s", __FILE__,        \
                       __LINE__, cudaGetErrorString(e__));                          \
  } while (0)

#define BDRC_REQUIRE(ctx, cond, ...)                                                \
  do {                                                                              \
    if (!(cond)) return bdrc_fail((ctx), BDRC_ERR_ARG, __VA_ARGS__);                \
  } while (0)

#define BDRC_TRY(expr)                                                              \
  do {                                                                              \
    int rc__ = (expr);                                                              \
    if (rc__ != 0) return rc__;                                                     \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// One bilinear sample in the operation order of torch's upsample_bilinear2d (horizontal lerps, then the
// vertical one), with the FMA contraction spelled out so that every kernel that upsamples logits -- the fused
// mask kernels of a-3 and the full-resolution upsample -- produces the same bits for the same inputs.
__device__ __forceinline__ float bdrc_bilerp(float a00, float a01, float a10, float a11, float lx0, float lx1,
                                             float ly0, float ly1) {
  const float top = __fmaf_rn(lx1, a01, __fmul_rn(lx0, a00));
  const float bot = __fmaf_rn(lx1, a11, __fmul_rn(lx0, a10));
  return __fmaf_rn(ly1, bot, __fmul_rn(ly0, top));
}
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- profiling helpers (conv_tc.cu / nets) ---------------------------------------
void bdrc_prof_begin(bdrc_ctx* ctx, const char* name, double flops, cudaStream_t s);
void bdrc_prof_end(bdrc_ctx* ctx, cudaStream_t s);

// ---- convolution problem (NHWC bf16 activations, packed bf16 weights) -------------
// Packed weight layout: [Cout_pad][R][S][Cin_pad64] bf16 (K-major rows), bias fp32[Cout_pad].
struct ConvWeights {
  bf16* w = nullptr;      // device, packed
  float* b = nullptr;     // device, Cout_pad floats (zeros if the layer has no bias)
  int Cout = 0, Cout_pad = 0;
  int Cin = 0, Cin_pad = 0;   // Cin_pad = round_up(Cin, 64)
  int R = 1, S = 1;
};

struct ConvArgs {
  const bf16* in = nullptr;   // [N,Hin,Win,in_ld] (+in_coff)
  int N = 0, Hin = 0, Win = 0, in_ld = 0, in_coff = 0;
  int stride_h = 1, stride_w = 1, pad_h = 0, pad_w = 0, dil_h = 1, dil_w = 1;
  int Hout = 0, Wout = 0;
  const float* bias_override = nullptr;   // optional per-image bias [N][Cout_pad]
  int bias_img_stride = 0;
  const bf16* residual = nullptr;         // optional [N,Hout,Wout,res_ld]
  int res_ld = 0;
  void* out = nullptr;                    // bf16 or fp32 [N,Hout,Wout,out_ld] (+out_coff)
  int out_ld = 0, out_coff = 0;
  int out_fp32 = 0;
  int relu = 0;
  int pdl = 0;                            // programmatic dependent launch (short chains only, see conv_tc.cu)
};

int conv_tc_supported(const ConvWeights& w, const ConvArgs& a);
int conv_tc_launch(bdrc_ctx* ctx, const ConvWeights& w, const ConvArgs& a, cudaStream_t s);
int conv_simt_launch(bdrc_ctx* ctx, const ConvWeights& w, const ConvArgs& a, cudaStream_t s);
void conv_tc_cache_free(bdrc_ctx* ctx);
// dispatch: tcgen05 (ctx->conv_engine == 0).  A shape the tcgen05 engine cannot tile is an ERROR unless the
// context allows the SIMT engine (BDRC_ALLOW_SIMT=1 or bdrc_set_conv_engine(ctx, 1)); SIMT launches are counted.
int conv_launch(bdrc_ctx* ctx, const char* name, const ConvWeights& w, const ConvArgs& a,
                cudaStream_t s);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link); nullptr if absent
typedef CUresult (*BdrcEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
BdrcEncodeTiledFn bdrc_tma_encode_fn();

// pack fp32 OIHW host weights into device ConvWeights
int conv_pack_weights(bdrc_ctx* ctx, const float* w_oihw, const float* bias, int Cout, int Cin,
                      int R, int S, ConvWeights* out);
void conv_free_weights(ConvWeights* w);
