// Context management and the small C-ABI entry points that are not tied to one stage.
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"
#include "nn_kernels.cuh"

extern "C" {

const char* bdrc_version(void) { return "bdrc_b200 0.1 (sm_100a)"; }

int bdrc_create(int device, bdrc_ctx** out) {
  if (!out) return BDRC_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || device < 0 || device >= n) {
    // No CPU fallback by design: the product path needs a B200.
    return BDRC_ERR_CUDA;
  }
  if (cudaSetDevice(device) != cudaSuccess) return BDRC_ERR_CUDA;
  bdrc_ctx* ctx = new bdrc_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete ctx;
    return BDRC_ERR_CUDA;
  }
  ctx->sm_count = prop.multiProcessorCount;
  if (prop.major != 10) {
    // the kernels are sm_100a-only (tcgen05/TMEM); refuse to run elsewhere
    delete ctx;
    return BDRC_ERR_UNSUPPORTED;
  }
  if (cudaMalloc(&ctx->d_errflag, sizeof(int)) != cudaSuccess) {
    delete ctx;
    return BDRC_ERR_CUDA;
  }
  cudaMemset(ctx->d_errflag, 0, sizeof(int));
  if (const char* e = getenv("BDRC_ALLOW_SIMT")) ctx->allow_simt = e[0] == '1';
  if (const char* e = getenv("BDRC_NO_DEC_TAIL")) ctx->no_dec_tail = e[0] == '1';
  if (const char* e = getenv("BDRC_NO_WIDE_CONV")) ctx->no_wide_conv = e[0] == '1';
  *out = ctx;
  return BDRC_OK;
}

void bdrc_destroy(bdrc_ctx* ctx) {
  if (!ctx) return;
  for (auto e : ctx->prof_ev) cudaEventDestroy(e);
  if (ctx->d_errflag) cudaFree(ctx->d_errflag);
  conv_tc_cache_free(ctx);
  delete ctx;
}

const char* bdrc_last_error(bdrc_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }

uint64_t bdrc_launch_count(bdrc_ctx* ctx) { return ctx ? ctx->launches : 0; }

uint64_t bdrc_simt_launch_count(bdrc_ctx* ctx) { return ctx ? ctx->simt_launches : 0; }

uint64_t bdrc_conv_plans_built(bdrc_ctx* ctx) { return ctx ? ctx->tc_plans_built : 0; }

int bdrc_set_conv_engine(bdrc_ctx* ctx, int engine) {
  if (!ctx || engine < 0 || engine > 1) return BDRC_ERR_ARG;
  ctx->conv_engine = engine;
  return BDRC_OK;
}

// Host threads that wait for the GPU (stream / event synchronisation, blocking copies) either spin on a core
// (the driver's default when there are more CPUs than contexts) or sleep until the device interrupts.  The
// page -> text path keeps ~10 such threads per GPU (page stages, upload ring, mask hand-out): spinning burns the
// host cores that the geometry needs and that an 8-GPU box does not have (4 per GPU).  Sets the scheduling flag
// of the device's primary context, which torch shares.
int bdrc_set_blocking_sync(bdrc_ctx* ctx, int on) {
  if (!ctx) return BDRC_ERR_ARG;
  BDRC_CUDA(ctx, cudaSetDevice(ctx->device));
  BDRC_CUDA(ctx, cudaSetDeviceFlags(on ? cudaDeviceScheduleBlockingSync : cudaDeviceScheduleAuto));
  return BDRC_OK;
}

int bdrc_set_dec_tail(bdrc_ctx* ctx, int fused) {
  if (!ctx) return BDRC_ERR_ARG;
  ctx->no_dec_tail = !fused;
  return BDRC_OK;
}

int bdrc_set_wide_conv(bdrc_ctx* ctx, int on) {
  if (!ctx) return BDRC_ERR_ARG;
  ctx->no_wide_conv = !on;
  return BDRC_OK;
}

int bdrc_profile_enable(bdrc_ctx* ctx, int on) {
  if (!ctx) return BDRC_ERR_ARG;
  ctx->profile = on != 0;
  return BDRC_OK;
}

int bdrc_profile_reset(bdrc_ctx* ctx) {
  if (!ctx) return BDRC_ERR_ARG;
  ctx->prof_used = 0;
  ctx->prof_name.clear();
  ctx->prof_flops.clear();
  return BDRC_OK;
}

int bdrc_profile_count(bdrc_ctx* ctx) { return ctx ? (int)(ctx->prof_used / 2) : 0; }

int bdrc_profile_get(bdrc_ctx* ctx, int i, const char** name, float* ms, double* flops) {
  if (!ctx || i < 0 || (size_t)i * 2 + 1 >= ctx->prof_used + 1 || (size_t)i >= ctx->prof_name.size())
    return BDRC_ERR_ARG;
  cudaEvent_t e0 = ctx->prof_ev[2 * i], e1 = ctx->prof_ev[2 * i + 1];
  BDRC_CUDA(ctx, cudaEventSynchronize(e1));
  float t = 0.f;
  BDRC_CUDA(ctx, cudaEventElapsedTime(&t, e0, e1));
  if (name) *name = ctx->prof_name[i].c_str();
  if (ms) *ms = t;
  if (flops) *flops = ctx->prof_flops[i];
  return BDRC_OK;
}

// A few persistent host threads for bdrc_host_gather: starting std::threads per call costs 50-100 us each inside
// a CUDA process (measured: it ate the whole gain of the parallel copy).  Jobs are index ranges [0, T); the
// caller runs index 0 itself.  One job at a time (calls from several Python threads queue on `call_`).
namespace {
class HostPool {
 public:
  static constexpr int kMax = 7;
  static HostPool& get() {
    static HostPool* p = new HostPool();   // never destroyed: the workers are detached and die with the process
    return *p;
  }
  void run(int T, const std::function<void(int)>& fn) {
    std::lock_guard<std::mutex> call(call_);
    {
      std::lock_guard<std::mutex> lk(m_);
      fn_ = &fn;
      total_ = T;
      next_ = 1;
      pending_ = T - 1;
      ++generation_;
    }
    cv_.notify_all();
    fn(0);
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  HostPool() {
    for (int i = 0; i < kMax; ++i) std::thread([this] { loop(); }).detach();
  }
  void loop() {
    uint64_t seen = 0;
    std::unique_lock<std::mutex> lk(m_);
    for (;;) {
      cv_.wait(lk, [&] { return generation_ != seen; });
      seen = generation_;
      while (fn_ && next_ < total_) {
        const int idx = next_++;
        const std::function<void(int)>* f = fn_;
        lk.unlock();
        (*f)(idx);
        lk.lock();
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  std::mutex m_, call_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int total_ = 0, next_ = 0, pending_ = 0;
  uint64_t generation_ = 0;
};
}  // namespace

// Host-side staging: n ragged byte ranges -> one (pinned) buffer, copied by a few native threads.
// One Python thread packing 64 line crops (19 MB) with numpy runs at memcpy speed of one core (~2 ms);
// this call releases the GIL once and splits the bytes evenly, so OCRInference.run_batch's upload
// is bound by the PCIe copy instead (BDRC/inference.py:336-347 builds its batch line by line).
int bdrc_host_gather(const void* const* src, const int64_t* nbytes, const int64_t* dst_off, int n, void* dst,
                     int threads) {
  if (n < 0 || (n > 0 && (!src || !nbytes || !dst_off || !dst))) return BDRC_ERR_ARG;
  int64_t total = 0;
  for (int i = 0; i < n; ++i) {
    if (nbytes[i] < 0 || dst_off[i] < 0 || (nbytes[i] > 0 && !src[i])) return BDRC_ERR_ARG;
    total += nbytes[i];
  }
  if (total == 0) return BDRC_OK;
  const int T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(threads, HostPool::kMax + 1), total / (1 << 20)));
  auto work = [&](int t) {
    // thread t copies bytes [lo, hi) of the concatenated stream
    const int64_t lo = total * t / T, hi = total * (t + 1) / T;
    int64_t pos = 0;
    for (int i = 0; i < n && pos < hi; ++i) {
      const int64_t a = std::max(lo, pos), b = std::min(hi, pos + nbytes[i]);
      if (b > a)
        std::memcpy(static_cast<uint8_t*>(dst) + dst_off[i] + (a - pos), static_cast<const uint8_t*>(src[i]) + (a - pos),
                    (size_t)(b - a));
      pos += nbytes[i];
    }
  };
  if (T == 1) {
    work(0);
    return BDRC_OK;
  }
  HostPool::get().run(T, work);
  return BDRC_OK;
}

// n independent host -> host copies (src[i] -> dst[i], nbytes[i]) split evenly over a few native threads:
// the hand-out of result masks as fresh caller-owned arrays (Detection.predict_batch, copy=True).
int bdrc_host_copy(const void* const* src, void* const* dst, const int64_t* nbytes, int n, int threads) {
  if (n < 0 || (n > 0 && (!src || !dst || !nbytes))) return BDRC_ERR_ARG;
  int64_t total = 0;
  for (int i = 0; i < n; ++i) {
    if (nbytes[i] < 0 || (nbytes[i] > 0 && (!src[i] || !dst[i]))) return BDRC_ERR_ARG;
    total += nbytes[i];
  }
  if (total == 0) return BDRC_OK;
  const int T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(threads, HostPool::kMax + 1), total / (1 << 20)));
  auto work = [&](int t) {
    const int64_t lo = total * t / T, hi = total * (t + 1) / T;
    int64_t pos = 0;
    for (int i = 0; i < n && pos < hi; ++i) {
      const int64_t a = std::max(lo, pos), b = std::min(hi, pos + nbytes[i]);
      if (b > a)
        std::memcpy(static_cast<uint8_t*>(dst[i]) + (a - pos), static_cast<const uint8_t*>(src[i]) + (a - pos), (size_t)(b - a));
      pos += nbytes[i];
    }
  };
  if (T == 1) {
    work(0);
    return BDRC_OK;
  }
  HostPool::get().run(T, work);
  return BDRC_OK;
}

int bdrc_conv2d_test(bdrc_ctx* ctx, const void* x, int N, int H, int W, int Cin,
                     const float* w_host, const float* bias_host, int Cout, int R, int S,
                     int stride_h, int stride_w, int pad_h, int pad_w, int dil_h, int dil_w,
                     const void* residual, int relu, int out_fp32, void* y, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  ConvWeights cw;
  BDRC_TRY(conv_pack_weights(ctx, w_host, bias_host, Cout, Cin, R, S, &cw));
  ConvArgs a;
  a.in = reinterpret_cast<const bf16*>(x);
  a.N = N; a.Hin = H; a.Win = W; a.in_ld = Cin; a.in_coff = 0;
  a.stride_h = stride_h; a.stride_w = stride_w; a.pad_h = pad_h; a.pad_w = pad_w;
  a.dil_h = dil_h; a.dil_w = dil_w;
  a.Hout = (H + 2 * pad_h - dil_h * (R - 1) - 1) / stride_h + 1;
  a.Wout = (W + 2 * pad_w - dil_w * (S - 1) - 1) / stride_w + 1;
  a.residual = reinterpret_cast<const bf16*>(residual);
  a.res_ld = Cout;
  a.out = y; a.out_ld = Cout; a.out_coff = 0; a.out_fp32 = out_fp32; a.relu = relu;
  int rc = conv_launch(ctx, "conv2d_test", cw, a, (cudaStream_t)stream);
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  conv_free_weights(&cw);
  if (rc) return rc;
  if (e != cudaSuccess) return bdrc_fail(ctx, BDRC_ERR_CUDA, "conv2d_test: %s", cudaGetErrorString(e));
  return BDRC_OK;
}

int bdrc_depthwise3x3_test(bdrc_ctx* ctx, const void* x, int N, int H, int W, int C, int dil,
                           const float* w9c_host, void* y, void* stream) {
  if (!ctx) return BDRC_ERR_ARG;
  float* w = nullptr;
  BDRC_CUDA(ctx, cudaMalloc(&w, (size_t)9 * C * sizeof(float)));
  cudaMemcpy(w, w9c_host, (size_t)9 * C * sizeof(float), cudaMemcpyHostToDevice);
  int rc = depthwise3x3_launch(ctx, reinterpret_cast<const bf16*>(x), C, 0, N, H, W, C, dil, w,
                               reinterpret_cast<bf16*>(y), C, 0, (cudaStream_t)stream);
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(w);
  if (rc) return rc;
  if (e != cudaSuccess) return bdrc_fail(ctx, BDRC_ERR_CUDA, "depthwise3x3_test: %s", cudaGetErrorString(e));
  return BDRC_OK;
}

}  // extern "C"
