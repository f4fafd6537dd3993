// SIMT convolution (debug cross-check engine; opt-in catch-all for shapes the tcgen05 engine does
// not tile), weight packing, conv dispatch and the per-launch profiling hooks.
#include <cstdarg>

#include "common.cuh"

int bdrc_fail(bdrc_ctx* ctx, int code, const char* fmt, ...) {
  if (ctx) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
    va_end(ap);
  }
  return code;
}

void bdrc_prof_begin(bdrc_ctx* ctx, const char* name, double flops, cudaStream_t s) {
  if (!ctx->profile) return;
  if (ctx->prof_used + 2 > ctx->prof_ev.size()) {
    for (int i = 0; i < 2; ++i) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ctx->prof_ev.push_back(e);
    }
  }
  ctx->prof_name.push_back(name);
  ctx->prof_flops.push_back(flops);
  cudaEventRecord(ctx->prof_ev[ctx->prof_used], s);
}

void bdrc_prof_end(bdrc_ctx* ctx, cudaStream_t s) {
  if (!ctx->profile) return;
  cudaEventRecord(ctx->prof_ev[ctx->prof_used + 1], s);
  ctx->prof_used += 2;
}

namespace {

__global__ void __launch_bounds__(64)
conv_simt_kernel(const bf16* __restrict__ in, int N, int Hin, int Win, int in_ld, int in_coff,
                 const bf16* __restrict__ w, int Cin, int Cin_pad, int R, int S, int Cout,
                 int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w, int Hout,
                 int Wout, const float* __restrict__ bias, int bias_img_stride,
                 const bf16* residual, int res_ld, void* out, int out_ld,
                 int out_coff, int out_fp32, int relu) {
  const long long M = (long long)N * Hout * Wout;
  const long long pix = (long long)blockIdx.x * 64 + threadIdx.x;
  const int co0 = blockIdx.y * 16;
  if (pix >= M) return;
  const int ox = (int)(pix % Wout);
  const int oy = (int)((pix / Wout) % Hout);
  const int n = (int)(pix / ((long long)Wout * Hout));
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  const size_t wrow = (size_t)R * S * Cin_pad;
  for (int r = 0; r < R; ++r) {
    const int iy = oy * stride_h - pad_h + r * dil_h;
    if (iy < 0 || iy >= Hin) continue;
    for (int s = 0; s < S; ++s) {
      const int ix = ox * stride_w - pad_w + s * dil_w;
      if (ix < 0 || ix >= Win) continue;
      const bf16* ip = in + (((size_t)n * Hin + iy) * Win + ix) * in_ld + in_coff;
      const bf16* wp = w + (size_t)co0 * wrow + (size_t)(r * S + s) * Cin_pad;
      int c = 0;
      for (; c + 8 <= Cin; c += 8) {
        const uint4 xv = *reinterpret_cast<const uint4*>(ip + c);
        const __nv_bfloat162* xh = reinterpret_cast<const __nv_bfloat162*>(&xv);
        float xf[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          xf[2 * q] = __low2float(xh[q]);
          xf[2 * q + 1] = __high2float(xh[q]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wp + (size_t)j * wrow + c));
          const __nv_bfloat162* wh = reinterpret_cast<const __nv_bfloat162*>(&wv);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            acc[j] = fmaf(xf[2 * q], __low2float(wh[q]), acc[j]);
            acc[j] = fmaf(xf[2 * q + 1], __high2float(wh[q]), acc[j]);
          }
        }
      }
      for (; c < Cin; ++c) {
        const float x = __bfloat162float(ip[c]);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          acc[j] = fmaf(x, __bfloat162float(wp[(size_t)j * wrow + c]), acc[j]);
      }
    }
  }
  const float* b = bias + (size_t)n * bias_img_stride + co0;
  for (int j = 0; j < 16; ++j) {
    const int co = co0 + j;
    if (co >= Cout) break;
    float v = acc[j] + b[j];
    if (residual) v += __bfloat162float(residual[(size_t)pix * res_ld + co]);
    if (relu) v = fmaxf(v, 0.f);
    if (out_fp32)
      reinterpret_cast<float*>(out)[(size_t)pix * out_ld + out_coff + co] = v;
    else
      reinterpret_cast<bf16*>(out)[(size_t)pix * out_ld + out_coff + co] = __float2bfloat16(v);
  }
}

inline uint16_t f32_to_bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

}  // namespace

int conv_simt_launch(bdrc_ctx* ctx, const ConvWeights& w, const ConvArgs& a, cudaStream_t s) {
  BDRC_REQUIRE(ctx, a.in_ld % 8 == 0 && a.in_coff % 8 == 0, "conv_simt: channel stride must be a multiple of 8");
  const long long M = (long long)a.N * a.Hout * a.Wout;
  dim3 grid((unsigned)((M + 63) / 64), (unsigned)ceil_div(w.Cout, 16));
  conv_simt_kernel<<<grid, 64, 0, s>>>(
      a.in, a.N, a.Hin, a.Win, a.in_ld, a.in_coff, w.w, w.Cin, w.Cin_pad, w.R, w.S, w.Cout,
      a.stride_h, a.stride_w, a.pad_h, a.pad_w, a.dil_h, a.dil_w, a.Hout, a.Wout,
      a.bias_override ? a.bias_override : w.b, a.bias_override ? a.bias_img_stride : 0, a.residual,
      a.res_ld, a.out, a.out_ld, a.out_coff, a.out_fp32, a.relu);
  BDRC_CHECK_LAUNCH(ctx);
  return 0;
}

int conv_launch(bdrc_ctx* ctx, const char* name, const ConvWeights& w, const ConvArgs& a,
                cudaStream_t s) {
  const double flops = 2.0 * a.N * a.Hout * a.Wout * (double)w.Cout * w.R * w.S * w.Cin;
  char label[160];
  if (ctx->profile) {
    snprintf(label, sizeof(label), "%s %d->%d @%dx%d k%dx%d s%d d%d", name, w.Cin, w.Cout, a.Hout, a.Wout, w.R, w.S,
             a.stride_w, a.dil_w);
    name = label;
  }
  int rc;
  if (ctx->conv_engine == 0 && conv_tc_supported(w, a)) {
    bdrc_prof_begin(ctx, name, flops, s);
    rc = conv_tc_launch(ctx, w, a, s);
  } else {
    // never silently: a model configuration that falls off the tcgen05 tiling would run ~50x slower
    if (ctx->conv_engine == 0 && !ctx->allow_simt)
      return bdrc_fail(ctx, BDRC_ERR_UNSUPPORTED,
                       "conv '%s' %d->%d k%dx%d s%dx%d on %dx%d (ld %d, off %d) has no tcgen05 tiling; "
                       "BDRC_ALLOW_SIMT=1 runs it on the (slow) SIMT engine",
                       name, w.Cin, w.Cout, w.R, w.S, a.stride_h, a.stride_w, a.Hout, a.Wout, a.in_ld, a.in_coff);
    ctx->simt_launches++;
    bdrc_prof_begin(ctx, name, flops, s);
    rc = conv_simt_launch(ctx, w, a, s);
  }
  bdrc_prof_end(ctx, s);
  return rc;
}

int conv_pack_weights(bdrc_ctx* ctx, const float* w_oihw, const float* bias, int Cout, int Cin,
                      int R, int S, ConvWeights* out) {
  out->Cout = Cout;
  out->Cin = Cin;
  out->R = R;
  out->S = S;
  out->Cout_pad = (Cout + 15) / 16 * 16;
  out->Cin_pad = (Cin + 63) / 64 * 64;
  const size_t K = (size_t)R * S * out->Cin_pad;
  std::vector<uint16_t> packed((size_t)out->Cout_pad * K, 0);
  for (int o = 0; o < Cout; ++o)
    for (int c = 0; c < Cin; ++c)
      for (int r = 0; r < R; ++r)
        for (int s = 0; s < S; ++s)
          packed[(size_t)o * K + (size_t)(r * S + s) * out->Cin_pad + c] =
              f32_to_bf16_rne(w_oihw[(((size_t)o * Cin + c) * R + r) * S + s]);
  std::vector<float> b(out->Cout_pad, 0.f);
  if (bias)
    for (int o = 0; o < Cout; ++o) b[o] = bias[o];
  BDRC_CUDA(ctx, cudaMalloc(&out->w, packed.size() * 2));
  BDRC_CUDA(ctx, cudaMalloc(&out->b, b.size() * 4));
  BDRC_CUDA(ctx, cudaMemcpy(out->w, packed.data(), packed.size() * 2, cudaMemcpyHostToDevice));
  BDRC_CUDA(ctx, cudaMemcpy(out->b, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  return 0;
}

void conv_free_weights(ConvWeights* w) {
  if (w->w) cudaFree(w->w);
  if (w->b) cudaFree(w->b);
  w->w = nullptr;
  w->b = nullptr;
}
