// Weight container parsing and the two network forward passes, assembled from the implicit-GEMM
// conv engine (conv_tc.cu) and the memory-bound kernels (nn_kernels.cu).
//
//   SEG : DeepLabV3+/ResNet-34, node list PyTorch-LineInference.ipynb:L233-L469 (BN folded)
//   OCR : Easter2 as declared in SURVEY.md 8(a)-8 (oracle/nets.py:easter2_forward is the twin)
#include <cstdlib>
#include <map>

#include "nn_kernels.cuh"

namespace {

struct BlobTensor {
  std::vector<int> dims;
  const float* data = nullptr;
  size_t numel = 0;
};

// "BDRCW1" container: magic[8], u32 n, then per tensor {u16 name_len, name, u8 ndim,
// u32 dims[ndim], u64 offset, u64 nbytes}, then (64-byte aligned) fp32 little-endian data.
bool parse_blob(const uint8_t* p, size_t n, std::map<std::string, BlobTensor>* out, std::string* err) {
  auto need = [&](size_t off, size_t k) { return off + k <= n; };
  if (n < 12 || memcmp(p, "BDRCW1\0\0", 8) != 0) { *err = "bad magic"; return false; }
  uint32_t cnt;
  memcpy(&cnt, p + 8, 4);
  size_t off = 12;
  struct Ent { std::string name; std::vector<int> dims; uint64_t o, b; };
  std::vector<Ent> ents;
  for (uint32_t i = 0; i < cnt; ++i) {
    if (!need(off, 2)) { *err = "truncated header"; return false; }
    uint16_t nl; memcpy(&nl, p + off, 2); off += 2;
    if (!need(off, nl + 1)) { *err = "truncated header"; return false; }
    Ent e; e.name.assign(reinterpret_cast<const char*>(p + off), nl); off += nl;
    const uint8_t nd = p[off++];
    if (nd > 8 || !need(off, 4 * nd + 16)) { *err = "truncated header"; return false; }
    for (int d = 0; d < nd; ++d) { uint32_t v; memcpy(&v, p + off, 4); off += 4; e.dims.push_back((int)v); }
    memcpy(&e.o, p + off, 8); off += 8;
    memcpy(&e.b, p + off, 8); off += 8;
    ents.push_back(e);
  }
  const size_t data0 = align_up(off, 64);
  for (auto& e : ents) {
    size_t numel = 1;
    for (int d : e.dims) numel *= (size_t)d;
    if (e.b != numel * 4 || !need(data0 + e.o, e.b)) { *err = "tensor '" + e.name + "' out of range"; return false; }
    BlobTensor t; t.dims = e.dims; t.numel = numel;
    t.data = reinterpret_cast<const float*>(p + data0 + e.o);
    (*out)[e.name] = t;
  }
  return true;
}

struct DwWeights { float* w9c = nullptr; int C = 0; };
struct FcWeights { float* w = nullptr; float* b = nullptr; int out = 0, in = 0; };

}  // namespace

struct bdrc_model {
  int kind = 0;
  int n_out = 0;
  bdrc_ctx* ctx = nullptr;
  std::map<std::string, ConvWeights> conv;
  std::map<std::string, DwWeights> dw;
  std::map<std::string, FcWeights> fc;
  std::vector<void*> owned;       // extra device allocations
  // seg
  float* stem_w3 = nullptr;       // [49*3][64]
  float* stem_w1 = nullptr;       // [49][64]   (3 identical input channels folded)
  float* stem_b = nullptr;
  bf16* stem_wb = nullptr;        // [64][64] bf16, k = 8 * row + col: folded weights for the tcgen05 stem
  float* proj_tail = nullptr;     // [256][256] fp32: projection columns of the pooled branch
  float* pool_w = nullptr;        // [256][512]
  float* pool_b = nullptr;
  float* proj_b = nullptr;
  float* head_wf = nullptr;       // [C][256] fp32 holding the bf16-rounded head weights (fused decoder tail)
  float* head_bf = nullptr;       // [C]
  std::vector<float> tail_bias_h, head_w_h, head_b_h;   // host copies: kernel parameters of the fused decoder tail
  int ocr_in_h = 0;
};

namespace {

int dev_copy(bdrc_ctx* ctx, bdrc_model* m, const std::vector<float>& h, float** out) {
  BDRC_CUDA(ctx, cudaMalloc(out, h.size() * sizeof(float)));
  m->owned.push_back(*out);
  BDRC_CUDA(ctx, cudaMemcpy(*out, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

struct Loader {
  bdrc_ctx* ctx;
  bdrc_model* m;
  std::map<std::string, BlobTensor>& t;

  const BlobTensor* get(const std::string& name, std::initializer_list<int> dims) {
    auto it = t.find(name);
    if (it == t.end()) { bdrc_fail(ctx, BDRC_ERR_WEIGHTS, "missing tensor '%s'", name.c_str()); return nullptr; }
    if (it->second.dims != std::vector<int>(dims)) {
      bdrc_fail(ctx, BDRC_ERR_WEIGHTS, "tensor '%s' has an unexpected shape", name.c_str());
      return nullptr;
    }
    return &it->second;
  }
  // 2-D conv [O][I][R][S]; 1-D conv [O][I][K] is loaded as R = 1, S = K
  int conv(const std::string& name, int O, int I, int R, int S, bool one_d = false, bool has_bias = true) {
    const BlobTensor* w = one_d ? get(name + ".weight", {O, I, S}) : get(name + ".weight", {O, I, R, S});
    if (!w) return BDRC_ERR_WEIGHTS;
    const float* b = nullptr;
    if (has_bias) {
      const BlobTensor* bt = get(name + ".bias", {O});
      if (!bt) return BDRC_ERR_WEIGHTS;
      b = bt->data;
    }
    ConvWeights cw;
    BDRC_TRY(conv_pack_weights(ctx, w->data, b, O, I, R, S, &cw));
    m->conv[name] = cw;
    return 0;
  }
  int depthwise(const std::string& name, int C) {
    const BlobTensor* w = get(name + ".weight", {C, 1, 3, 3});
    if (!w) return BDRC_ERR_WEIGHTS;
    std::vector<float> h((size_t)9 * C);
    for (int c = 0; c < C; ++c)
      for (int k = 0; k < 9; ++k) h[(size_t)k * C + c] = w->data[(size_t)c * 9 + k];
    DwWeights d; d.C = C;
    BDRC_TRY(dev_copy(ctx, m, h, &d.w9c));
    m->dw[name] = d;
    return 0;
  }
  int linear(const std::string& name, int O, int I) {
    const BlobTensor* w = get(name + ".weight", {O, I});
    const BlobTensor* b = get(name + ".bias", {O});
    if (!w || !b) return BDRC_ERR_WEIGHTS;
    FcWeights f; f.out = O; f.in = I;
    BDRC_TRY(dev_copy(ctx, m, std::vector<float>(w->data, w->data + w->numel), &f.w));
    BDRC_TRY(dev_copy(ctx, m, std::vector<float>(b->data, b->data + b->numel), &f.b));
    m->fc[name] = f;
    return 0;
  }
};

const int R34_CH[4] = {64, 128, 256, 512};
const int R34_BLOCKS[4] = {3, 4, 6, 3};
const int R34_STRIDE[4] = {1, 2, 2, 1};
const int R34_DIL[4] = {1, 1, 1, 2};

int load_seg(Loader& L, int n_classes) {
  bdrc_ctx* ctx = L.ctx;
  bdrc_model* m = L.m;
  // stem: accept 3-channel weights (ipynb:L131); keep a folded 1-channel copy for binarised tiles
  const BlobTensor* sw = L.get("encoder.conv1.weight", {64, 3, 7, 7});
  const BlobTensor* sb = L.get("encoder.conv1.bias", {64});
  if (!sw || !sb) return BDRC_ERR_WEIGHTS;
  std::vector<float> w3((size_t)147 * 64), w1((size_t)49 * 64, 0.f);
  for (int o = 0; o < 64; ++o)
    for (int c = 0; c < 3; ++c)
      for (int k = 0; k < 49; ++k) {
        const float v = sw->data[((size_t)o * 3 + c) * 49 + k];
        w3[((size_t)c * 49 + k) * 64 + o] = v;
        w1[(size_t)k * 64 + o] += v;
      }
  BDRC_TRY(dev_copy(ctx, m, w3, &m->stem_w3));
  BDRC_TRY(dev_copy(ctx, m, w1, &m->stem_w1));
  BDRC_TRY(dev_copy(ctx, m, std::vector<float>(sb->data, sb->data + 64), &m->stem_b));
  {
    std::vector<bf16> wb((size_t)64 * 64, __float2bfloat16(0.f));
    for (int o = 0; o < 64; ++o)
      for (int r = 0; r < 7; ++r)
        for (int c = 0; c < 7; ++c) wb[(size_t)o * 64 + r * 8 + c] = __float2bfloat16(w1[(size_t)(r * 7 + c) * 64 + o]);
    BDRC_CUDA(ctx, cudaMalloc(&m->stem_wb, wb.size() * sizeof(bf16)));
    m->owned.push_back(m->stem_wb);
    BDRC_CUDA(ctx, cudaMemcpy(m->stem_wb, wb.data(), wb.size() * sizeof(bf16), cudaMemcpyHostToDevice));
  }
  int cin = 64;
  for (int li = 0; li < 4; ++li) {
    const int ch = R34_CH[li];
    for (int b = 0; b < R34_BLOCKS[li]; ++b) {
      const std::string p = "encoder.layer" + std::to_string(li + 1) + "." + std::to_string(b);
      BDRC_TRY(L.conv(p + ".conv1", ch, b == 0 ? cin : ch, 3, 3));
      BDRC_TRY(L.conv(p + ".conv2", ch, ch, 3, 3));
      if (b == 0 && (cin != ch || R34_STRIDE[li] != 1)) BDRC_TRY(L.conv(p + ".downsample.0", ch, cin, 1, 1));
    }
    cin = ch;
  }
  BDRC_TRY(L.conv("decoder.aspp.0.convs.0.0", 256, 512, 1, 1));
  for (int i = 1; i <= 3; ++i) {
    const std::string p = "decoder.aspp.0.convs." + std::to_string(i);
    BDRC_TRY(L.depthwise(p + ".0.0", 512));
    BDRC_TRY(L.conv(p + ".0.1", 256, 512, 1, 1));
  }
  // pooled branch (1x1 on the global mean) and the projection, split into the 1024 "spatial"
  // input columns (GEMM) and the 256 pooled columns (folded into a per-tile bias)
  {
    const BlobTensor* pw = L.get("decoder.aspp.0.convs.4.1.weight", {256, 512, 1, 1});
    const BlobTensor* pb = L.get("decoder.aspp.0.convs.4.1.bias", {256});
    const BlobTensor* jw = L.get("decoder.aspp.0.project.0.weight", {256, 1280, 1, 1});
    const BlobTensor* jb = L.get("decoder.aspp.0.project.0.bias", {256});
    if (!pw || !pb || !jw || !jb) return BDRC_ERR_WEIGHTS;
    BDRC_TRY(dev_copy(ctx, m, std::vector<float>(pw->data, pw->data + pw->numel), &m->pool_w));
    BDRC_TRY(dev_copy(ctx, m, std::vector<float>(pb->data, pb->data + 256), &m->pool_b));
    BDRC_TRY(dev_copy(ctx, m, std::vector<float>(jb->data, jb->data + 256), &m->proj_b));
    std::vector<float> head((size_t)256 * 1024), tail((size_t)256 * 256);
    for (int o = 0; o < 256; ++o) {
      memcpy(&head[(size_t)o * 1024], jw->data + (size_t)o * 1280, 1024 * sizeof(float));
      memcpy(&tail[(size_t)o * 256], jw->data + (size_t)o * 1280 + 1024, 256 * sizeof(float));
    }
    BDRC_TRY(dev_copy(ctx, m, tail, &m->proj_tail));
    ConvWeights cw;
    BDRC_TRY(conv_pack_weights(ctx, head.data(), nullptr, 256, 1024, 1, 1, &cw));
    m->conv["decoder.aspp.0.project.0"] = cw;
  }
  BDRC_TRY(L.depthwise("decoder.aspp.1.0", 256));
  BDRC_TRY(L.conv("decoder.aspp.1.1", 256, 256, 1, 1));
  BDRC_TRY(L.conv("decoder.block1.0", 48, 64, 1, 1));
  BDRC_TRY(L.depthwise("decoder.block2.0.0", 304));
  BDRC_TRY(L.conv("decoder.block2.0.1", 256, 304, 1, 1));
  BDRC_TRY(L.conv("segmentation_head.0", n_classes, 256, 1, 1));
  {
    const BlobTensor* hw = L.get("segmentation_head.0.weight", {n_classes, 256, 1, 1});
    const BlobTensor* hb = L.get("segmentation_head.0.bias", {n_classes});
    if (!hw || !hb) return BDRC_ERR_WEIGHTS;
    std::vector<float> wf(hw->numel);
    for (size_t i = 0; i < hw->numel; ++i) wf[i] = __bfloat162float(__float2bfloat16(hw->data[i]));   // as conv_pack_weights stores them
    BDRC_TRY(dev_copy(ctx, m, wf, &m->head_wf));
    BDRC_TRY(dev_copy(ctx, m, std::vector<float>(hb->data, hb->data + n_classes), &m->head_bf));
    const BlobTensor* tb = L.get("decoder.block2.0.1.bias", {256});
    if (!tb) return BDRC_ERR_WEIGHTS;
    m->tail_bias_h.assign(tb->data, tb->data + 256);
    m->head_w_h = wf;
    m->head_b_h.assign(hb->data, hb->data + n_classes);
  }
  return 0;
}

int load_ocr(Loader& L, int vocab) {
  auto it = L.t.find("conv1.weight");
  if (it == L.t.end() || it->second.dims.size() != 3)
    return bdrc_fail(L.ctx, BDRC_ERR_WEIGHTS, "missing tensor 'conv1.weight'");
  const int in_h = it->second.dims[1];
  L.m->ocr_in_h = in_h;
  BDRC_TRY(L.conv("conv1", 128, in_h, 1, 3, true));
  BDRC_TRY(L.conv("conv2", 128, 128, 1, 3, true));
  int cin = 128;
  const int ks[3] = {5, 7, 9};
  for (int u = 0; u < 3; ++u) {
    const std::string p = "unit" + std::to_string(u + 1);
    // res = conv1x1(old) + conv1x1(data) as ONE GEMM: in unit 1 old == data (weights add), later the two
    // inputs are adjacent channel slices [old | data] of one buffer (weights concatenate along K)
    {
      const BlobTensor* wo = L.get(p + ".res_old.weight", {256, cin, 1});
      const BlobTensor* wd = L.get(p + ".res_data.weight", {256, cin, 1});
      const BlobTensor* bo = L.get(p + ".res_old.bias", {256});
      const BlobTensor* bd = L.get(p + ".res_data.bias", {256});
      if (!wo || !wd || !bo || !bd) return BDRC_ERR_WEIGHTS;
      const int kin = (u == 0) ? cin : 2 * cin;
      std::vector<float> wc((size_t)256 * kin), bc(256);
      for (int o = 0; o < 256; ++o) {
        bc[o] = bo->data[o] + bd->data[o];
        for (int c = 0; c < cin; ++c) {
          if (u == 0) {
            wc[(size_t)o * kin + c] = wo->data[(size_t)o * cin + c] + wd->data[(size_t)o * cin + c];
          } else {
            wc[(size_t)o * kin + c] = wo->data[(size_t)o * cin + c];
            wc[(size_t)o * kin + cin + c] = wd->data[(size_t)o * cin + c];
          }
        }
      }
      ConvWeights cw;
      BDRC_TRY(conv_pack_weights(L.ctx, wc.data(), bc.data(), 256, kin, 1, 1, &cw));
      L.m->conv[p + ".res"] = cw;
    }
    BDRC_TRY(L.conv(p + ".conv_a", 256, cin, 1, ks[u], true));
    BDRC_TRY(L.conv(p + ".conv_b", 256, 256, 1, ks[u], true));
    BDRC_TRY(L.conv(p + ".conv_c", 256, 256, 1, ks[u], true));
    BDRC_TRY(L.linear(p + ".se_fc1", 32, 256));
    BDRC_TRY(L.linear(p + ".se_fc2", 256, 32));
    cin = 256;
  }
  BDRC_TRY(L.conv("conv3", 512, 256, 1, 11, true));
  BDRC_TRY(L.conv("conv4", 512, 512, 1, 1, true));
  BDRC_TRY(L.conv("head", vocab, 512, 1, 1, true));
  return 0;
}

// bump allocator over the caller's workspace
struct Arena {
  uint8_t* base; size_t cap, off = 0; bool ok = true;
  template <class T> T* take(size_t n) {
    const size_t bytes = align_up(n * sizeof(T), 1024);
    if (off + bytes > cap) { ok = false; return nullptr; }
    T* p = reinterpret_cast<T*>(base + off);
    off += bytes;
    return p;
  }
};

// Tiles per forward chunk.  Multiples of 37 give every ResNet stage a whole number of waves on 148 SMs / 74 CTA
// pairs (a 32x32 stage has 8 pixel tiles per image), and larger chunks amortise the ~57 launches of a forward:
// measured (round 1 kernels) 91 us per tile at 48 tiles, 80 at 74, 77 at 148; round 2, the 384 tiles of a C2 step on one
// box: 25.92 / 26.05 ms in chunks of 148 (148 + 148 + 88), 25.79 at 192, 25.60 / 25.22 in ONE chunk of 384 (a third of
// the launches, one partial last wave per layer instead of three; 19 GB of activations instead of 7.4).
constexpr int SEG_CHUNK_DEFAULT = 384;
static int seg_chunk() {   // BDRC_SEG_CHUNK overrides (sweeps: tools/seg_chunk_sweep.sh); read once
  static const int v = [] {
    const char* e = getenv("BDRC_SEG_CHUNK");
    const int x = e ? atoi(e) : 0;
    return (x >= 1 && x <= 1024) ? x : SEG_CHUNK_DEFAULT;
  }();
  return v;
}
constexpr int SEG_P = 512;

size_t seg_bytes_per_tile(int C) {
  size_t b = 0;
  auto add = [&](size_t elems, size_t sz) { b += align_up(elems * sz, 1024); };
  add((size_t)256 * 256 * 64, 2);           // stem
  for (int i = 0; i < 3; ++i) add((size_t)128 * 128 * 64, 2);   // rotating block buffers
  add((size_t)128 * 128 * 64, 2);           // skip (layer1 output)
  add((size_t)32 * 32 * 1024, 2);           // ASPP concat
  add((size_t)3 * 32 * 32 * 512, 2);        // depthwise scratch (one per ASPP rate)
  add((size_t)256, 4);                      // per-tile projection bias
  add((size_t)512, 4);                      // per-tile channel means (image-pooling branch)
  add((size_t)32 * 32 * 256, 2);            // projection out
  add((size_t)32 * 32 * 256, 2);            // aspp.1 depthwise
  add((size_t)32 * 32 * 256, 2);            // aspp.1 pointwise
  add((size_t)128 * 128 * 304, 2);          // decoder concat
  add((size_t)128 * 128 * 304, 2);          // block2 depthwise
  add((size_t)128 * 128 * 256, 2);          // block2 pointwise
  add((size_t)128 * 128 * (size_t)((C + 3) / 4 * 4), 4);   // head, low-res fp32 NHWC
  return b + 16 * 1024;
}

ConvArgs mk(const bf16* in, int N, int H, int W, int ld, int coff, int k, int stride, int dil, void* out,
            int out_ld, int out_coff, int relu) {
  ConvArgs a;
  a.in = in; a.N = N; a.Hin = H; a.Win = W; a.in_ld = ld; a.in_coff = coff;
  a.stride_h = a.stride_w = stride;
  a.dil_h = a.dil_w = dil;
  a.pad_h = a.pad_w = (k / 2) * dil;
  a.Hout = H / stride; a.Wout = W / stride;
  a.out = out; a.out_ld = out_ld; a.out_coff = out_coff; a.relu = relu;
  return a;
}

int seg_forward_chunk(bdrc_ctx* ctx, bdrc_model* m, const void* tiles, int n, int cin, float* logits_full,
                      float* logits_lowres, uint8_t* ws, size_t ws_bytes, cudaStream_t s) {
  const int C = m->n_out;
  Arena A{ws, ws_bytes};
  bf16* stem = A.take<bf16>((size_t)n * 256 * 256 * 64);
  bf16* buf[3];
  for (int i = 0; i < 3; ++i) buf[i] = A.take<bf16>((size_t)n * 128 * 128 * 64);
  bf16* skip = A.take<bf16>((size_t)n * 128 * 128 * 64);
  bf16* cat = A.take<bf16>((size_t)n * 32 * 32 * 1024);
  bf16* dws = A.take<bf16>((size_t)3 * n * 32 * 32 * 512);
  float* pbias = A.take<float>((size_t)n * 256);
  float* gmean = A.take<float>((size_t)n * 512);
  bf16* proj = A.take<bf16>((size_t)n * 32 * 32 * 256);
  bf16* a1dw = A.take<bf16>((size_t)n * 32 * 32 * 256);
  bf16* a1pw = A.take<bf16>((size_t)n * 32 * 32 * 256);
  bf16* dcat = A.take<bf16>((size_t)n * 128 * 128 * 304);
  bf16* d2dw = A.take<bf16>((size_t)n * 128 * 128 * 304);
  bf16* d2pw = A.take<bf16>((size_t)n * 128 * 128 * 256);
  const int Cq = (C + 3) / 4 * 4;
  float* low = logits_lowres ? logits_lowres : A.take<float>((size_t)n * 128 * 128 * Cq);
  const int low_ld = C;
  if (!A.ok) return bdrc_fail(ctx, BDRC_ERR_ARG, "seg_forward: workspace too small");

  // stem + maxpool (ipynb:L233-L238)
  if (cin == 1 && ctx->conv_engine == 0) {
    // binarised tiles: tcgen05 stem fused with the max-pool, the 256x256x64 activation stays on chip
    bdrc_prof_begin(ctx, "stem7x7+maxpool", 2.0 * n * 256 * 256 * 64 * 49 * 3, s);
    BDRC_TRY(stem_tc_pool_launch(ctx, reinterpret_cast<const uint8_t*>(tiles), n, m->stem_wb, m->stem_b, buf[0], s));
    bdrc_prof_end(ctx, s);
  } else {
    bdrc_prof_begin(ctx, "stem7x7", 2.0 * n * 256 * 256 * 64 * 49 * 3, s);
    BDRC_TRY(stem_conv_launch(ctx, tiles, cin, cin == 1 ? 1.f : 255.f, n, SEG_P, SEG_P,
                              cin == 1 ? m->stem_w1 : m->stem_w3, m->stem_b, stem, s));
    bdrc_prof_end(ctx, s);
    bdrc_prof_begin(ctx, "maxpool", 0, s);
    BDRC_TRY(maxpool3x3s2_launch(ctx, stem, n, 256, 256, 64, buf[0], s));
    bdrc_prof_end(ctx, s);
  }

  // ResNet-34 body (ipynb:L239-L404).  x walks through three rotating buffers; the output of
  // layer1 is written to `skip`, which stays untouched until the decoder reads it.
  bf16* x = buf[0];
  auto pick = [&](const bf16* a, const bf16* b) -> bf16* {
    for (int i = 0; i < 3; ++i)
      if (buf[i] != a && buf[i] != b) return buf[i];
    return nullptr;
  };
  int size = 128, cinl = 64;
  for (int li = 0; li < 4; ++li) {
    const int ch = R34_CH[li], dil = R34_DIL[li];
    for (int b = 0; b < R34_BLOCKS[li]; ++b) {
      const std::string p = "encoder.layer" + std::to_string(li + 1) + "." + std::to_string(b);
      const int st = b == 0 ? R34_STRIDE[li] : 1;
      const int ci = b == 0 ? cinl : ch;
      const int so = size / st;
      bf16* y = pick(x, nullptr);
      const bool last_l1 = (li == 0 && b == R34_BLOCKS[0] - 1);
      bf16* outb = last_l1 ? skip : pick(x, y);
      BDRC_TRY(conv_launch(ctx, "r34.conv3x3", m->conv[p + ".conv1"],
                           mk(x, n, size, size, ci, 0, 3, st, dil, y, ch, 0, 1), s));
      const bf16* idt = x;
      if (m->conv.count(p + ".downsample.0")) {
        // projection shortcut goes to the output buffer first and is then read as the residual
        // by the same threads that overwrite it (each element is read before it is written)
        BDRC_TRY(conv_launch(ctx, "r34.down1x1", m->conv[p + ".downsample.0"],
                             mk(x, n, size, size, ci, 0, 1, st, 1, outb, ch, 0, 0), s));
        idt = outb;
      }
      ConvArgs a2 = mk(y, n, so, so, ch, 0, 3, 1, dil, outb, ch, 0, 1);
      a2.residual = idt; a2.res_ld = ch;
      BDRC_TRY(conv_launch(ctx, "r34.conv3x3", m->conv[p + ".conv2"], a2, s));
      x = outb;
      size = so;
    }
    cinl = ch;
  }
  // x: [n,32,32,512]

  // ASPP (ipynb:L405-L444): four spatial branches write channel slices of `cat`
  BDRC_TRY(conv_launch(ctx, "aspp.1x1", m->conv["decoder.aspp.0.convs.0.0"],
                       mk(x, n, 32, 32, 512, 0, 1, 1, 1, cat, 1024, 0, 1), s));
  const int rates[3] = {12, 24, 36};
  {
    const size_t br = (size_t)n * 32 * 32 * 512;
    const float* dww[3];
    bf16* dwo[3];
    for (int i = 0; i < 3; ++i) {
      dww[i] = m->dw["decoder.aspp.0.convs." + std::to_string(i + 1) + ".0.0"].w9c;
      dwo[i] = dws + i * br;
    }
    bdrc_prof_begin(ctx, "aspp.dw3x3", 2.0 * n * 32 * 32 * 512 * 9 * 3, s);
    BDRC_TRY(aspp_dw3_launch(ctx, x, n, 512, rates, dww, dwo, gmean, s));
    bdrc_prof_end(ctx, s);
    for (int i = 0; i < 3; ++i)
      BDRC_TRY(conv_launch(ctx, "aspp.1x1", m->conv["decoder.aspp.0.convs." + std::to_string(i + 1) + ".0.1"],
                           mk(dwo[i], n, 32, 32, 512, 0, 1, 1, 1, cat, 1024, 256 * (i + 1), 1), s));
  }
  bdrc_prof_begin(ctx, "aspp.pool", 2.0 * n * (256.0 * 512 + 256 * 256), s);
  BDRC_TRY(aspp_pool_bias_launch(ctx, x, n, 32 * 32, 512, m->pool_w, m->pool_b, 256, m->proj_tail,
                                 m->proj_b, 256, 256, pbias, gmean, s));
  bdrc_prof_end(ctx, s);
  {
    ConvArgs a = mk(cat, n, 32, 32, 1024, 0, 1, 1, 1, proj, 256, 0, 1);
    a.bias_override = pbias; a.bias_img_stride = 256;
    BDRC_TRY(conv_launch(ctx, "aspp.project", m->conv["decoder.aspp.0.project.0"], a, s));
  }
  bdrc_prof_begin(ctx, "dec.dw3x3", 2.0 * n * 32 * 32 * 256 * 9, s);
  BDRC_TRY(depthwise3x3_launch(ctx, proj, 256, 0, n, 32, 32, 256, 1, m->dw["decoder.aspp.1.0"].w9c, a1dw, 256, 0, s));
  bdrc_prof_end(ctx, s);
  BDRC_TRY(conv_launch(ctx, "dec.1x1", m->conv["decoder.aspp.1.1"],
                       mk(a1dw, n, 32, 32, 256, 0, 1, 1, 1, a1pw, 256, 0, 1), s));
  // decoder (ipynb:L445-L464): x4 upsample into dcat[:, :256], skip 1x1 into dcat[:, 256:304]
  bdrc_prof_begin(ctx, "dec.upsample", 0, s);
  BDRC_TRY(upsample_bilinear_ac_launch(ctx, a1pw, n, 32, 32, 256, 4, dcat, 304, 0, s));
  bdrc_prof_end(ctx, s);
  BDRC_TRY(conv_launch(ctx, "dec.skip1x1", m->conv["decoder.block1.0"],
                       mk(skip, n, 128, 128, 64, 0, 1, 1, 1, dcat, 304, 256, 1), s));
  // depthwise 3x3 (304) -> 1x1 304 -> 256 + ReLU -> head 1x1 256 -> C (ipynb:L455-L466); fp32 NHWC low-res logits.
  if (ctx->conv_engine == 0 && !ctx->no_dec_tail && dec_tail_supported(304, 256, C, 128, 128)) {
    // one kernel: the 304- and 256-channel activations of this resolution never reach HBM (dec_tail.cu)
    bdrc_prof_begin(ctx, "dec.tail dw3x3+1x1+head", 2.0 * n * 128 * 128 * (304.0 * 9 + 256.0 * 304 + (double)C * 256), s);
    BDRC_TRY(dec_tail_launch(ctx, dcat, n, 128, 128, 304, m->dw["decoder.block2.0.0"].w9c, m->conv["decoder.block2.0.1"],
                             m->tail_bias_h.data(), m->head_w_h.data(), m->head_b_h.data(), C, low, s));
    bdrc_prof_end(ctx, s);
  } else {
    bdrc_prof_begin(ctx, "dec.dw3x3", 2.0 * n * 128 * 128 * 304 * 9, s);
    BDRC_TRY(depthwise3x3_launch(ctx, dcat, 304, 0, n, 128, 128, 304, 1, m->dw["decoder.block2.0.0"].w9c, d2dw, 304, 0, s));
    bdrc_prof_end(ctx, s);
    BDRC_TRY(conv_launch(ctx, "dec.1x1", m->conv["decoder.block2.0.1"],
                         mk(d2dw, n, 128, 128, 304, 0, 1, 1, 1, d2pw, 256, 0, 1), s));
    ConvArgs a = mk(d2pw, n, 128, 128, 256, 0, 1, 1, 1, low, low_ld, 0, 0);
    a.out_fp32 = 1;
    BDRC_TRY(conv_launch(ctx, "head1x1", m->conv["segmentation_head.0"], a, s));
  }
  if (logits_full) {
    // final x4 align_corners upsample (ipynb:L467-L469) to NCHW fp32
    bdrc_prof_begin(ctx, "head.upsample", 0, s);
    BDRC_TRY(upsample_logits_launch(ctx, low, n, 128, 128, C, logits_full, s));
    bdrc_prof_end(ctx, s);
  }
  return 0;
}

int seg_forward_impl(bdrc_ctx* ctx, bdrc_model* m, const void* tiles, int n_tiles, int cin, float* logits,
                     float* logits_lowres, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (!ctx || !m) return BDRC_ERR_ARG;
  BDRC_REQUIRE(ctx, m->kind == BDRC_MODEL_SEG_DLV3P_R34, "seg_forward: not a segmentation model");
  BDRC_REQUIRE(ctx, tiles && (logits || logits_lowres) && n_tiles >= 0 && (cin == 1 || cin == 3),
               "seg_forward: bad arguments");
  BDRC_REQUIRE(ctx, ws && ws_bytes >= bdrc_seg_workspace_bytes(m, n_tiles), "seg_forward: workspace too small");
  const int C = m->n_out;
  const size_t tile_in = (size_t)SEG_P * SEG_P * (cin == 1 ? 1 : 3 * sizeof(float));
  const int SEG_CHUNK = seg_chunk();
  for (int t0 = 0; t0 < n_tiles; t0 += SEG_CHUNK) {
    const int n = n_tiles - t0 < SEG_CHUNK ? n_tiles - t0 : SEG_CHUNK;
    BDRC_TRY(seg_forward_chunk(
        ctx, m, reinterpret_cast<const uint8_t*>(tiles) + (size_t)t0 * tile_in, n, cin,
        logits ? logits + (size_t)t0 * C * SEG_P * SEG_P : nullptr,
        logits_lowres ? logits_lowres + (size_t)t0 * 128 * 128 * C : nullptr,
        reinterpret_cast<uint8_t*>(ws), ws_bytes, s));
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Easter2 forward.  Activations are bf16 time-major [L][T][C].
// ---------------------------------------------------------------------------------------------
constexpr int OCR_CPAD = 128;   // conv1 input channels (image rows) padded for the stride-2 view

ConvArgs mk1d(const bf16* in, int L, int T, int ld, int k, int stride, int dil, void* out, int out_ld, int relu) {
  ConvArgs a;
  a.in = in; a.N = L; a.Hin = 1; a.Win = T; a.in_ld = ld; a.in_coff = 0;
  a.stride_h = 1; a.stride_w = stride; a.dil_h = 1; a.dil_w = dil;
  a.pad_h = 0; a.pad_w = (k / 2) * dil;
  a.Hout = 1; a.Wout = (T + 2 * a.pad_w - dil * (k - 1) - 1) / stride + 1;
  a.out = out; a.out_ld = out_ld; a.out_coff = 0; a.relu = relu;
  a.pdl = 1;   // the Easter2 chain is short (26 launches): PDL gains 3 % at 64 lines (conv_tc.cu)
  return a;
}

size_t ocr_bytes_per_line(int Wt, int vocab) {
  const size_t T2 = Wt / 2, T = Wt / 4;
  size_t b = 0;
  auto add = [&](size_t elems, size_t sz) { b += align_up(elems * sz, 1024); };
  add(T2 * 128, 2);
  for (int i = 0; i < 5; ++i) add(T * 512, 2);
  add((size_t)256 * (se_fused_chunks() + 1) + 32, 4);
  (void)vocab;
  return b + 8 * 1024;
}

int ocr_forward_impl(bdrc_ctx* ctx, bdrc_model* m, const bf16* lines, int L, int Wt, float* logits, void* ws,
                     size_t ws_bytes, cudaStream_t s) {
  if (!ctx || !m) return BDRC_ERR_ARG;
  BDRC_REQUIRE(ctx, m->kind == BDRC_MODEL_OCR_EASTER2, "ocr_forward: not an OCR model");
  BDRC_REQUIRE(ctx, lines && logits && L >= 0 && Wt > 0 && Wt % 4 == 0, "ocr_forward: bad arguments (Wt %% 4)");
  BDRC_REQUIRE(ctx, m->ocr_in_h <= OCR_CPAD, "ocr_forward: input height %d > %d", m->ocr_in_h, OCR_CPAD);
  if (L == 0) return 0;
  BDRC_REQUIRE(ctx, ws && ws_bytes >= bdrc_ocr_workspace_bytes(m, L, Wt), "ocr_forward: workspace too small");
  const int T2 = Wt / 2, T = Wt / 4, V = m->n_out;
  Arena A{reinterpret_cast<uint8_t*>(ws), ws_bytes};
  bf16* c1 = A.take<bf16>((size_t)L * T2 * 128);
  bf16* b0 = A.take<bf16>((size_t)L * T * 512);
  bf16* b1 = A.take<bf16>((size_t)L * T * 512);
  bf16* b2 = A.take<bf16>((size_t)L * T * 512);
  bf16* b3 = A.take<bf16>((size_t)L * T * 512);
  bf16* b4 = A.take<bf16>((size_t)L * T * 512);
  float* gate = A.take<float>(se_fused_scratch_floats(L, 256));   // SE partial sums, gates, ticket counters
  if (!A.ok) return bdrc_fail(ctx, BDRC_ERR_ARG, "ocr_forward: workspace too small");

  BDRC_TRY(se_fused_reset(ctx, gate, L, 256, s));
  // conv1 sees the image rows as channels; they were zero-padded to OCR_CPAD by the prep kernel,
  // the packed weights' Cin_pad covers it (extra K columns are zero)
  ConvWeights w1 = m->conv["conv1"];
  w1.Cin = OCR_CPAD;   // activation view has OCR_CPAD channels per time step
  BDRC_TRY(conv_launch(ctx, "ocr.conv1", w1, mk1d(lines, L, Wt, OCR_CPAD, 3, 2, 1, c1, 128, 1), s));
  BDRC_TRY(conv_launch(ctx, "ocr.conv2", m->conv["conv2"], mk1d(c1, L, T2, 128, 3, 2, 1, b0, 128, 1), s));
  // Each unit writes its outputs into one 512-channel "pair" buffer P = [res | data]; the next unit's
  // residual GEMM reads the whole pair (K = 512), its k-tap convs read the data slice.
  bf16* pairs[2] = {b1, b2};
  const bf16* prev = b0;       // unit input: c2 (128 ch) for unit 1, else the previous pair buffer
  int prev_ld = 128, data_off = 0, cin = 128;
  const int ks[3] = {5, 7, 9};
  bf16* da = b3;
  bf16* db = b4;
  for (int u = 0; u < 3; ++u) {
    const std::string p = "unit" + std::to_string(u + 1);
    bf16* P = pairs[u & 1];
    {
      ConvArgs a = mk1d(prev, L, T, prev_ld, 1, 1, 1, P, 512, 0);
      BDRC_TRY(conv_launch(ctx, "ocr.res1x1", m->conv[p + ".res"], a, s));
    }
    {
      ConvArgs a = mk1d(prev, L, T, prev_ld, ks[u], 1, 1, da, 256, 1);
      a.in_coff = data_off;
      BDRC_TRY(conv_launch(ctx, "ocr.convk", m->conv[p + ".conv_a"], a, s));
    }
    BDRC_TRY(conv_launch(ctx, "ocr.convk", m->conv[p + ".conv_b"], mk1d(da, L, T, 256, ks[u], 1, 1, db, 256, 1), s));
    BDRC_TRY(conv_launch(ctx, "ocr.convk", m->conv[p + ".conv_c"], mk1d(db, L, T, 256, ks[u], 1, 1, da, 256, 0), s));
    const FcWeights& f1 = m->fc[p + ".se_fc1"];
    const FcWeights& f2 = m->fc[p + ".se_fc2"];
    bdrc_prof_begin(ctx, "ocr.se", 0, s);
    BDRC_TRY(se_fused_launch(ctx, P, 512, da, L, T, 256, f1.w, f1.b, 32, f2.w, f2.b, P + 256, 512, gate, /*pdl=*/1, s));
    bdrc_prof_end(ctx, s);
    prev = P; prev_ld = 512; data_off = 256; cin = 256;
  }
  (void)cin;
  bf16* fr[2] = {pairs[1], b3};   // pairs[1] (unit 2's pair) is dead once unit 3 has consumed it
  {
    ConvArgs a = mk1d(prev, L, T, 512, 11, 1, 2, fr[0], 512, 1);
    a.in_coff = 256;
    BDRC_TRY(conv_launch(ctx, "ocr.conv3", m->conv["conv3"], a, s));
  }
  BDRC_TRY(conv_launch(ctx, "ocr.conv4", m->conv["conv4"], mk1d(fr[0], L, T, 512, 1, 1, 1, fr[1], 512, 1), s));
  {
    ConvArgs a = mk1d(fr[1], L, T, 512, 1, 1, 1, logits, V, 0);
    a.out_fp32 = 1;
    BDRC_TRY(conv_launch(ctx, "ocr.head", m->conv["head"], a, s));
  }
  return 0;
}

}  // namespace

extern "C" {

int bdrc_load_weights(bdrc_ctx* ctx, int model_kind, const void* blob_host, size_t blob_bytes, int n_out,
                      bdrc_model** out) {
  if (!ctx || !out) return BDRC_ERR_ARG;
  *out = nullptr;
  BDRC_REQUIRE(ctx, blob_host && n_out >= 1 && n_out <= 4096, "load_weights: bad arguments");
  std::map<std::string, BlobTensor> t;
  std::string err;
  if (!parse_blob(reinterpret_cast<const uint8_t*>(blob_host), blob_bytes, &t, &err))
    return bdrc_fail(ctx, BDRC_ERR_WEIGHTS, "weight container: %s", err.c_str());
  bdrc_model* m = new bdrc_model();
  m->kind = model_kind;
  m->n_out = n_out;
  m->ctx = ctx;
  Loader L{ctx, m, t};
  int rc;
  if (model_kind == BDRC_MODEL_SEG_DLV3P_R34) {
    rc = n_out <= 32 ? load_seg(L, n_out) : bdrc_fail(ctx, BDRC_ERR_ARG, "at most 32 classes");
  } else if (model_kind == BDRC_MODEL_OCR_EASTER2) {
    rc = load_ocr(L, n_out);
  } else {
    rc = bdrc_fail(ctx, BDRC_ERR_ARG, "unknown model kind %d", model_kind);
  }
  if (rc) {
    bdrc_free_model(m);
    return rc;
  }
  *out = m;
  return 0;
}

void bdrc_free_model(bdrc_model* m) {
  if (!m) return;
  for (auto& kv : m->conv) conv_free_weights(&kv.second);
  for (void* p : m->owned) cudaFree(p);
  delete m;
}

size_t bdrc_seg_workspace_bytes(bdrc_model* m, int n_tiles) {
  if (!m || n_tiles <= 0) return 0;
  const int n = n_tiles < seg_chunk() ? n_tiles : seg_chunk();
  return seg_bytes_per_tile(m->n_out) * (size_t)n;
}

int bdrc_seg_forward(bdrc_ctx* ctx, bdrc_model* m, const void* tiles, int n_tiles, int cin, float* logits,
                     void* workspace, size_t workspace_bytes, void* stream) {
  return seg_forward_impl(ctx, m, tiles, n_tiles, cin, logits, nullptr, workspace, workspace_bytes,
                          (cudaStream_t)stream);
}

int bdrc_seg_forward_lowres(bdrc_ctx* ctx, bdrc_model* m, const void* tiles, int n_tiles, int cin,
                            float* logits_lowres, void* workspace, size_t workspace_bytes, void* stream) {
  return seg_forward_impl(ctx, m, tiles, n_tiles, cin, nullptr, logits_lowres, workspace, workspace_bytes,
                          (cudaStream_t)stream);
}

size_t bdrc_ocr_workspace_bytes(bdrc_model* m, int L, int Wt) {
  if (!m || L <= 0 || Wt <= 0) return 0;
  return ocr_bytes_per_line(Wt, m->n_out) * (size_t)L;
}

int bdrc_ocr_forward(bdrc_ctx* ctx, bdrc_model* m, const void* lines_bf16, int L, int Wt, float* logits,
                     void* workspace, size_t workspace_bytes, void* stream) {
  return ocr_forward_impl(ctx, m, reinterpret_cast<const bf16*>(lines_bf16), L, Wt, logits, workspace,
                          workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
