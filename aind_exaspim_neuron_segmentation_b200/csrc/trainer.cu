// Trainer (trainer.h): schedules one training step layer by layer.
//   forward : reference unet3d.py:77-105 with BatchNorm3d in training mode (unet3d.py:144,147)
//   backward: what loss.backward() does for that graph (train.py:139-147)
#include "trainer.h"

#include <stdlib.h>

#include <algorithm>

namespace exa {

namespace {

void free_dev(void* p) {
  if (p) cudaFree(p);
}

const char* kPrefix[9] = {
    "inc.double_conv",
    "down1.maxpool_conv.1.double_conv", "down2.maxpool_conv.1.double_conv",
    "down3.maxpool_conv.1.double_conv", "down4.maxpool_conv.1.double_conv",
    "up1.conv.double_conv", "up2.conv.double_conv", "up3.conv.double_conv", "up4.conv.double_conv"};

size_t align_up(size_t n, size_t a) { return (n + a - 1) / a * a; }

TView plain(const Act& a) {
  TView v;
  v.a = a;
  v.enc = false;
  return v;
}
TView encoded(const Act& a) {
  TView v;
  v.a = a;
  v.enc = true;
  return v;
}
// channels [c0, c0 + n) of a concat-shaped tensor
Act channel_slice(const Act& a, int c0, int n) {
  Act s = a;
  s.coff = a.coff + c0;
  s.C = n;
  return s;
}

}  // namespace

#define EXA_LAUNCH(cat, expr)   \
  do {                         \
    Scope _sc(this, (cat), s); \
    EXA_TRY(expr);             \
  } while (0)

Trainer::Scope::Scope(Trainer* tr, int cat, cudaStream_t st) : t(tr), s(st) {
  ++t->launches;
  if (!t->prof_on_) return;
  ProfRec r;
  r.cat = cat;
  if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess) return;
  cudaEventRecord(r.start, s);
  stop = r.stop;
  t->prof_.push_back(r);
}
Trainer::Scope::~Scope() {
  if (stop) cudaEventRecord(stop, s);
}

Status Trainer::profile_begin() {
  EXA_CUDA(cudaSetDevice(device_));
  for (auto& r : prof_) {
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  prof_.clear();
  prof_on_ = true;
  return Status::OK();
}

Status Trainer::profile_end(double* ms_by_cat, int64_t* launches_by_cat, int n) {
  EXA_CHECK(ms_by_cat && launches_by_cat && n >= CAT_COUNT, "train_profile_end: need >= 12 slots");
  EXA_CUDA(cudaSetDevice(device_));
  prof_on_ = false;
  for (int i = 0; i < n; ++i) {
    ms_by_cat[i] = 0.0;
    launches_by_cat[i] = 0;
  }
  EXA_CUDA(cudaDeviceSynchronize());
  for (auto& r : prof_) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) {
      ms_by_cat[r.cat] += ms;
      ++launches_by_cat[r.cat];
    }
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  prof_.clear();
  return Status::OK();
}

Trainer::~Trainer() {
  cudaSetDevice(device_);
  for (auto& r : prof_) {
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  free_dev(ws_);
  free_dev(small_);
}

Status Trainer::init() {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    return Status::Err(std::string("no CUDA device available (") + cudaGetErrorString(e) +
                       "); this library has no CPU fallback");
  }
  EXA_CHECK(device_ >= 0 && device_ < count, "device index out of range");
  EXA_CHECK(precision_ == EXA_PRECISION_BF16 || precision_ == EXA_PRECISION_FP32,
            "unknown precision");
  EXA_CUDA(cudaSetDevice(device_));
  cudaDeviceProp prop;
  EXA_CUDA(cudaGetDeviceProperties(&prop, device_));
  EXA_CHECK(prop.major == 10, "this library is built for sm_100a (B200) only; found sm_" +
                                  std::to_string(prop.major) + std::to_string(prop.minor));
  num_sms_ = prop.multiProcessorCount;
  return Status::OK();
}

Status Trainer::bind(const char* name, float* dev_ptr, const int64_t* shape, int ndim) {
  EXA_CHECK(name && dev_ptr && ndim >= 0 && ndim <= 5, "train_bind: bad arguments");
  Bound b;
  b.ptr = dev_ptr;
  for (int i = 0; i < ndim; ++i) b.shape.push_back(shape[i]);
  bound_[name] = b;
  resolved_ = false;
  forward_valid_ = false;
  return Status::OK();
}

// Checks the bound entries against the reference layout (unet3d.py:56-75 with trilinear=True)
// and lays out the gradient slots in state_dict order of the parameters.
Status Trainer::resolve() {
  if (resolved_) return Status::OK();
  EXA_CUDA(cudaSetDevice(device_));
  auto get = [&](const std::string& key, std::vector<int64_t> shape, float** out) -> Status {
    auto it = bound_.find(key);
    EXA_CHECK(it != bound_.end(), "missing state_dict entry: " + key);
    EXA_CHECK(it->second.shape == shape, "wrong shape for state_dict entry: " + key);
    *out = it->second.ptr;
    return Status::OK();
  };
  {
    auto it = bound_.find("inc.double_conv.0.weight");
    EXA_CHECK(it != bound_.end(), "missing state_dict entry: inc.double_conv.0.weight");
    EXA_CHECK(it->second.shape.size() == 5 && it->second.shape[0] == 32,
              "training supports width_multiplier = 1 (32 channels in the first block)");
    EXA_CHECK(bound_.find("up1.up.weight") == bound_.end(),
              "training supports trilinear=True models only");
    for (int k = 0; k < 5; ++k) chan_[k] = 32 << k;
  }
  const int* c = chan_;
  const int io[9][3] = {{1, c[0], c[0]},          {c[0], c[1], c[1]},     {c[1], c[2], c[2]},
                        {c[2], c[3], c[3]},       {c[3], c[4] / 2, c[4] / 2},
                        {c[4], c[4] / 2, c[3] / 2}, {c[3], c[3] / 2, c[2] / 2},
                        {c[2], c[2] / 2, c[1] / 2}, {c[1], c[1] / 2, c[0]}};
  const int lvl[9] = {0, 1, 2, 3, 4, 3, 2, 1, 0};
  slots_.clear();
  int64_t off = 0;
  auto slot = [&](const std::string& key, int64_t n) {
    slots_[key] = {off, n};
    const int64_t o = off;
    off += (n + 3) / 4 * 4;  // 16-byte aligned slots
    return o;
  };
  for (int blk = 0; blk < 9; ++blk)
    for (int half = 0; half < 2; ++half) {
      Layer& L = layers_[2 * blk + half];
      L.conv_key = std::string(kPrefix[blk]) + (half ? ".3" : ".0");
      L.bn_key = std::string(kPrefix[blk]) + (half ? ".4" : ".1");
      L.cin = half ? io[blk][1] : io[blk][0];
      L.cout = half ? io[blk][2] : io[blk][1];
      L.lvl = lvl[blk];
      EXA_TRY(get(L.conv_key + ".weight", {L.cout, L.cin, 3, 3, 3}, &L.w));
      EXA_TRY(get(L.conv_key + ".bias", {L.cout}, &L.b));
      EXA_TRY(get(L.bn_key + ".weight", {L.cout}, &L.gamma));
      EXA_TRY(get(L.bn_key + ".bias", {L.cout}, &L.beta));
      EXA_TRY(get(L.bn_key + ".running_mean", {L.cout}, &L.rmean));
      EXA_TRY(get(L.bn_key + ".running_var", {L.cout}, &L.rvar));
      L.gw = slot(L.conv_key + ".weight", (int64_t)L.cout * L.cin * 27);
      L.gb = slot(L.conv_key + ".bias", L.cout);
      L.ggamma = slot(L.bn_key + ".weight", L.cout);
      L.gbeta = slot(L.bn_key + ".bias", L.cout);
    }
  {
    auto it = bound_.find("outc.conv.weight");
    EXA_CHECK(it != bound_.end(), "missing state_dict entry: outc.conv.weight");
    EXA_CHECK(it->second.shape.size() == 5 && it->second.shape[1] == c[0] &&
                  it->second.shape[0] >= 1 && it->second.shape[0] <= 8,
              "wrong shape for state_dict entry: outc.conv.weight");
    out_channels_ = (int)it->second.shape[0];
    EXA_TRY(get("outc.conv.weight", {out_channels_, c[0], 1, 1, 1}, &head_w_));
    EXA_TRY(get("outc.conv.bias", {out_channels_}, &head_b_));
    g_head_w_ = slot("outc.conv.weight", (int64_t)out_channels_ * c[0]);
    g_head_b_ = slot("outc.conv.bias", out_channels_);
  }
  grad_elems_ = off;

  // small device arena: per-layer statistics and packed weights
  const bool f32 = precision_ == EXA_PRECISION_FP32;
  const size_t wsz = f32 ? 4 : 2;
  size_t bytes = 0;
  auto take = [&](size_t n) {
    const size_t o = bytes;
    bytes += align_up(n, 256);
    return o;
  };
  struct Offs {
    size_t mean, rstd, scale, shift, coef, sums, w_fwd, w_fwd_zf, w_bwd, w_bwd_zf;
  } offs[18];
  for (int l = 0; l < 18; ++l) {
    const Layer& L = layers_[l];
    offs[l].mean = take(4 * L.cout);
    offs[l].rstd = take(4 * L.cout);
    offs[l].scale = take(4 * L.cout);
    offs[l].shift = take(4 * L.cout);
    offs[l].coef = take(4 * 3 * L.cout);
    offs[l].sums = take(8 * 2 * L.cout);
    const size_t n = (size_t)27 * L.cin * L.cout;
    if (l == 0) {
      // stem: Toeplitz band matrices (bf16) or the 16-input-channel form (fp32)
      offs[l].w_fwd = take(f32 ? (size_t)27 * 16 * L.cout * 4 : (size_t)9 * 128 * 16 * (L.cout / 32) * 2);
      offs[l].w_fwd_zf = offs[l].w_bwd = offs[l].w_bwd_zf = (size_t)-1;
    } else {
      offs[l].w_fwd = take(n * wsz);
      offs[l].w_bwd = take(n * wsz);
      offs[l].w_fwd_zf = f32 ? (size_t)-1 : take(n * wsz);
      offs[l].w_bwd_zf = f32 ? (size_t)-1 : take(n * wsz);
    }
  }
  const size_t off_zero = take(4 * 1024);
  const size_t off_head = take(8 * ((size_t)out_channels_ * c[0] + out_channels_ + 8));
  if (bytes > small_bytes_) {
    free_dev(small_);
    small_ = nullptr;
    small_bytes_ = 0;
    EXA_CUDA(cudaMalloc(&small_, bytes));
    small_bytes_ = bytes;
  }
  EXA_CUDA(cudaMemset(small_, 0, bytes));
  char* base = (char*)small_;
  auto at = [&](size_t o) -> void* { return o == (size_t)-1 ? nullptr : (void*)(base + o); };
  for (int l = 0; l < 18; ++l) {
    Layer& L = layers_[l];
    L.mean = (float*)at(offs[l].mean);
    L.rstd = (float*)at(offs[l].rstd);
    L.scale = (float*)at(offs[l].scale);
    L.shift = (float*)at(offs[l].shift);
    L.coef = (float*)at(offs[l].coef);
    L.sums = (double*)at(offs[l].sums);
    L.w_fwd = at(offs[l].w_fwd);
    L.w_fwd_zf = at(offs[l].w_fwd_zf);
    L.w_bwd = at(offs[l].w_bwd);
    L.w_bwd_zf = at(offs[l].w_bwd_zf);
  }
  zero_bias_ = (float*)at(off_zero);
  head_sums_ = (double*)at(off_head);
  ws_batch_ = 0;  // views are rebuilt on the next forward
  resolved_ = true;
  return Status::OK();
}

Status Trainer::grad_slot(const char* name, int64_t* offset, int64_t* numel) {
  EXA_CHECK(name && offset && numel, "train_grad_slot: null argument");
  EXA_TRY(resolve());
  auto it = slots_.find(name);
  EXA_CHECK(it != slots_.end(), std::string("not a parameter: ") + name);
  *offset = it->second.first;
  *numel = it->second.second;
  return Status::OK();
}

Status Trainer::grad_elems(int64_t* n) {
  EXA_CHECK(n, "train_grad_elems: null argument");
  EXA_TRY(resolve());
  *n = grad_elems_;
  return Status::OK();
}

// Every tensor of one step has its own buffer: the backward pass reads the conv inputs (weight
// gradients), the raw conv outputs (BatchNorm backward) and the activations (LeakyReLU mask,
// max-pool argmax), so nothing of the inference workspace's aliasing survives.
Status Trainer::ensure_workspace(int batch, int pz, int py, int px) {
  if (ws_ && batch == ws_batch_ && pz == ws_p_[0] && py == ws_p_[1] && px == ws_p_[2])
    return Status::OK();
  const bool f32 = precision_ == EXA_PRECISION_FP32;
  const size_t esz = f32 ? 4 : 2;
  const int* c = chan_;
  const int cb = c[4] / 2;
  const int m1 = c[4] / 2, o1 = c[3] / 2, m2 = c[3] / 2, o2 = c[2] / 2, m3 = c[2] / 2, o3 = c[1] / 2;
  size_t bytes = 0;
  struct Req {
    Act* act;
    int lvl, C;
    size_t off;
  };
  std::vector<Req> reqs;
  auto vox = [&](int lvl) { return (size_t)batch * (pz >> lvl) * (py >> lvl) * (px >> lvl); };
  auto want = [&](Act* a, int lvl, int C) {
    Req r{a, lvl, C, bytes};
    bytes += align_up(vox(lvl) * C * esz, 256);
    reqs.push_back(r);
  };
  // activations
  want(&a0_, 0, c[0]); want(&cat4_, 0, 2 * c[0]); want(&p1_, 1, c[0]); want(&d1a_, 1, c[1]);
  want(&cat3_, 1, 2 * c[1]); want(&p2_, 2, c[1]); want(&d2a_, 2, c[2]); want(&cat2_, 2, 2 * c[2]);
  want(&p3_, 3, c[2]); want(&d3a_, 3, c[3]); want(&cat1_, 3, 2 * c[3]); want(&p4_, 4, c[3]);
  want(&d4a_, 4, cb); want(&x5_, 4, cb); want(&u1a_, 3, m1); want(&u1_, 3, o1); want(&u2a_, 2, m2);
  want(&u2_, 2, o2); want(&u3a_, 1, m3); want(&u3_, 1, o3); want(&u4a_, 0, c[0]);
  want(&u4_, 0, c[0]);
  // raw conv outputs
  for (int l = 0; l < 18; ++l) want(&layers_[l].z, layers_[l].lvl, layers_[l].cout);
  // gradients
  want(&g_u4_, 0, c[0]); want(&g_u4a_, 0, c[0]); want(&g_cat4_, 0, 2 * c[0]); want(&g_u3_, 1, o3);
  want(&g_u3a_, 1, m3); want(&g_cat3_, 1, 2 * c[1]); want(&g_u2_, 2, o2); want(&g_u2a_, 2, m2);
  want(&g_cat2_, 2, 2 * c[2]); want(&g_u1_, 3, o1); want(&g_u1a_, 3, m1);
  want(&g_cat1_, 3, 2 * c[3]); want(&g_x5_, 4, cb); want(&g_d4a_, 4, cb); want(&g_p4_, 4, c[3]);
  want(&g_x4_, 3, c[3]); want(&g_d3a_, 3, c[3]); want(&g_p3_, 3, c[2]); want(&g_x3_, 2, c[2]);
  want(&g_d2a_, 2, c[2]); want(&g_p2_, 2, c[1]); want(&g_x2_, 1, c[1]); want(&g_d1a_, 1, c[1]);
  want(&g_p1_, 1, c[0]); want(&g_x1_, 0, c[0]); want(&g_a0_, 0, c[0]);
  // gradient of the raw conv output of the layer in flight (largest: level 0)
  size_t dz_elems = 0;
  for (int l = 0; l < 18; ++l) dz_elems = std::max(dz_elems, vox(layers_[l].lvl) * layers_[l].cout);
  const size_t off_dz = bytes;
  bytes += align_up(dz_elems * esz, 256);
  // stem input: bf16 (hi, lo) pairs on rows of px + 8 voxels, or 16 fp32 channels per voxel
  const size_t off_stem = bytes;
  bytes += align_up(f32 ? vox(0) * 16 * 4 : (size_t)batch * pz * py * (px + 8) * 2 * 2, 256);

  auto make = [&](char* base, const Req& r) {
    Act a;
    a.ptr = base + r.off;
    a.B = batch;
    a.D = pz >> r.lvl;
    a.H = py >> r.lvl;
    a.W = px >> r.lvl;
    a.C = r.C;
    a.cstride = r.C;
    a.coff = 0;
    a.fp32 = f32;
    return a;
  };
  // split-K partial sums of the weight gradients: needs the conv-input shapes, so build the
  // views against a null base first
  for (const Req& r : reqs) *r.act = make(nullptr, r);
  size_t partial_elems = 0;
  {
    const int in_lvl[18] = {0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 3, 3, 2, 2, 1, 1, 0, 0};
    for (int l = 1; l < 18; ++l) {
      Act x = a0_;
      x.D = pz >> in_lvl[l]; x.H = py >> in_lvl[l]; x.W = px >> in_lvl[l];
      x.C = layers_[l].cin;
      partial_elems = std::max(partial_elems, wgrad_partial_elems(x, layers_[l].cout, num_sms_));
    }
    Act dz0 = a0_;
    partial_elems = std::max(partial_elems,
                             (size_t)wgrad_stem_splits(dz0, num_sms_) * layers_[0].cout * 27);
  }
  const size_t off_partial = bytes;
  bytes += align_up(partial_elems * 4, 256);

  if (bytes > ws_bytes_) {
    free_dev(ws_);
    ws_ = nullptr;
    ws_bytes_ = 0;
    cudaError_t e = cudaMalloc(&ws_, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return Status::Err("cudaMalloc of the training workspace (" + std::to_string(bytes >> 20) +
                         " MiB) failed: " + cudaGetErrorString(e));
    }
    ws_bytes_ = bytes;
  }
  char* base = (char*)ws_;
  for (const Req& r : reqs) *r.act = make(base, r);
  {
    Req r{nullptr, 0, c[0], off_dz};
    dz_ = make(base, r);
  }
  stem_in_ = base + off_stem;
  partial_ = (float*)(base + off_partial);
  // views into the concat buffers (unet3d.py:288: cat([skip, upsampled]))
  x1_ = channel_slice(cat4_, 0, c[0]); up4_slot_ = channel_slice(cat4_, c[0], c[0]);
  x2_ = channel_slice(cat3_, 0, c[1]); up3_slot_ = channel_slice(cat3_, c[1], c[1]);
  x3_ = channel_slice(cat2_, 0, c[2]); up2_slot_ = channel_slice(cat2_, c[2], c[2]);
  x4_ = channel_slice(cat1_, 0, c[3]); up1_slot_ = channel_slice(cat1_, c[3], c[3]);
  // conv input / activation / input-gradient of every layer, in unet3d.py:64-74 order
  const Act none{};
  const Act* xin[18] = {&none, &a0_, &p1_, &d1a_, &p2_, &d2a_, &p3_, &d3a_, &p4_, &d4a_,
                        &cat1_, &u1a_, &cat2_, &u2a_, &cat3_, &u3a_, &cat4_, &u4a_};
  const Act* aout[18] = {&a0_, &x1_, &d1a_, &x2_, &d2a_, &x3_, &d3a_, &x4_, &d4a_, &x5_,
                         &u1a_, &u1_, &u2a_, &u2_, &u3a_, &u3_, &u4a_, &u4_};
  const Act* gx[18] = {&none, &g_a0_, &g_p1_, &g_d1a_, &g_p2_, &g_d2a_, &g_p3_, &g_d3a_, &g_p4_,
                       &g_d4a_, &g_cat1_, &g_u1a_, &g_cat2_, &g_u2a_, &g_cat3_, &g_u3a_, &g_cat4_,
                       &g_u4a_};
  for (int l = 0; l < 18; ++l) {
    layers_[l].x = *xin[l];
    layers_[l].a = *aout[l];
    layers_[l].gx = *gx[l];
  }
  ws_batch_ = batch;
  ws_p_[0] = pz; ws_p_[1] = py; ws_p_[2] = px;
  return Status::OK();
}

// 3x3x3 convolution + bias through the inference kernels (their epilogue's LeakyReLU makes the
// output an encoded tensor, train_kernels.h)
Status Trainer::conv_any(const Act& in, const Act& out, const void* w_plain, const void* w_zf,
                         const float* bias, cudaStream_t s) {
  if (precision_ == EXA_PRECISION_FP32)
    return launch_conv_fp32(in, out, (const float*)w_plain, bias, s);
  static const bool no_zf = getenv("EXA_TRAIN_NO_ZFOLD") != nullptr;  // A/B: generic kernel only
  if (!no_zf && w_zf && conv_zfold_supported(in, out.C, true))
    return launch_conv_zfold(in, out, (const __nv_bfloat16*)w_zf, bias, nullptr, nullptr, nullptr,
                             num_sms_, true, s);
  return launch_conv_umma(in, out, (const __nv_bfloat16*)w_plain, bias, nullptr, num_sms_, s);
}

Status Trainer::layer_forward(int l, const float* x, cudaStream_t s) {
  Layer& L = layers_[l];
  const bool f32 = precision_ == EXA_PRECISION_FP32;
  if (l == 0) {
    if (f32) {
      EXA_LAUNCH(CAT_PACK, launch_pack_stem_fp32(L.w, (float*)L.w_fwd, L.cout, s));
      EXA_LAUNCH(CAT_PACK, launch_expand_input16(x, (float*)stem_in_, L.z.voxels(), s));
      Act in = L.z;
      in.ptr = stem_in_;
      in.C = in.cstride = 16;
      EXA_LAUNCH(CAT_FPROP, launch_conv_fp32(in, L.z, (const float*)L.w_fwd, L.b, s));
    } else {
      const int groups = L.cout / 32;
      EXA_LAUNCH(CAT_PACK, launch_pack_stem_band(L.w, (__nv_bfloat16*)L.w_fwd, groups, s));
      PatchSource src;
      src.x = x;
      EXA_LAUNCH(CAT_FPROP, launch_stem_split(src, L.z.B, L.z.D, L.z.H, L.z.W, (__nv_bfloat16*)stem_in_, s));
      for (int g = 0; g < groups; ++g)
        EXA_LAUNCH(CAT_FPROP, launch_stem_tc((const __nv_bfloat16*)stem_in_,
                                  (const __nv_bfloat16*)L.w_fwd + (size_t)g * 9 * 128 * 16,
                                  L.b + g * 32, channel_slice(L.z, g * 32, 32), num_sms_, s));
    }
  } else {
    // forward operand [tap][Cout][Cin] (bf16; fp32: [tap][Cin][Cout]); the data gradient is the
    // same convolution with the taps flipped and the channel roles swapped
    EXA_LAUNCH(CAT_PACK, launch_pack_conv_weights(L.w, L.w_fwd, L.cout, L.cin, !f32, false, false, f32, s));
    EXA_LAUNCH(CAT_PACK, launch_pack_conv_weights(L.w, L.w_bwd, L.cout, L.cin, f32, true, false, f32, s));
    if (L.w_fwd_zf)
      EXA_LAUNCH(CAT_PACK, launch_pack_conv_weights(L.w, L.w_fwd_zf, L.cout, L.cin, true, false, true, false, s));
    if (L.w_bwd_zf)
      EXA_LAUNCH(CAT_PACK, launch_pack_conv_weights(L.w, L.w_bwd_zf, L.cout, L.cin, false, true, true, false, s));
    EXA_LAUNCH(CAT_FPROP, conv_any(L.x, L.z, L.w_fwd, L.w_fwd_zf, L.b, s));
  }
  const double count = (double)L.z.voxels();
  EXA_CUDA(cudaMemsetAsync(L.sums, 0, sizeof(double) * 2 * L.cout, s));
  EXA_LAUNCH(CAT_BN_FWD, launch_bn_stats(encoded(L.z), L.sums, s));
  EXA_LAUNCH(CAT_BN_FWD, launch_bn_finalize(L.sums, L.cout, count, L.gamma, L.beta, L.rmean, L.rvar, L.mean,
                                L.rstd, L.scale, L.shift, s));
  EXA_LAUNCH(CAT_BN_FWD, launch_bn_apply(encoded(L.z), L.scale, L.shift, L.a, s));
  return Status::OK();
}

Status Trainer::forward(const float* x, int batch, const int32_t patch[3], float* logits,
                        cudaStream_t s) {
  EXA_CUDA(cudaSetDevice(device_));
  EXA_CHECK(x && logits && batch > 0, "train_forward: bad arguments");
  for (int i = 0; i < 3; ++i)
    EXA_CHECK(patch[i] > 0 && patch[i] % 16 == 0,
              "train_forward: patch dims must be multiples of 16");
  EXA_CHECK((int64_t)batch * (patch[0] >> 4) * (patch[1] >> 4) * (patch[2] >> 4) > 1,
            "train_forward: BatchNorm in training mode needs more than one value per channel");
  EXA_TRY(resolve());
  forward_valid_ = false;
  EXA_TRY(ensure_workspace(batch, patch[0], patch[1], patch[2]));

  EXA_TRY(layer_forward(0, x, s));
  EXA_TRY(layer_forward(1, x, s));
  EXA_LAUNCH(CAT_MISC_FWD, launch_maxpool(x1_, p1_, s));
  EXA_TRY(layer_forward(2, x, s));
  EXA_TRY(layer_forward(3, x, s));
  EXA_LAUNCH(CAT_MISC_FWD, launch_maxpool(x2_, p2_, s));
  EXA_TRY(layer_forward(4, x, s));
  EXA_TRY(layer_forward(5, x, s));
  EXA_LAUNCH(CAT_MISC_FWD, launch_maxpool(x3_, p3_, s));
  EXA_TRY(layer_forward(6, x, s));
  EXA_TRY(layer_forward(7, x, s));
  EXA_LAUNCH(CAT_MISC_FWD, launch_maxpool(x4_, p4_, s));
  EXA_TRY(layer_forward(8, x, s));
  EXA_TRY(layer_forward(9, x, s));
  EXA_LAUNCH(CAT_MISC_FWD, launch_upsample(x5_, up1_slot_, nullptr, s));  // cat([x4, up(x5)]), unet3d.py:288
  EXA_TRY(layer_forward(10, x, s));
  EXA_TRY(layer_forward(11, x, s));
  EXA_LAUNCH(CAT_MISC_FWD, launch_upsample(u1_, up2_slot_, nullptr, s));
  EXA_TRY(layer_forward(12, x, s));
  EXA_TRY(layer_forward(13, x, s));
  EXA_LAUNCH(CAT_MISC_FWD, launch_upsample(u2_, up3_slot_, nullptr, s));
  EXA_TRY(layer_forward(14, x, s));
  EXA_TRY(layer_forward(15, x, s));
  EXA_LAUNCH(CAT_MISC_FWD, launch_upsample(u3_, up4_slot_, nullptr, s));
  EXA_TRY(layer_forward(16, x, s));
  EXA_TRY(layer_forward(17, x, s));
  HeadParams head;
  head.w = head_w_;
  head.b = head_b_;
  head.out = logits;
  head.C = out_channels_;
  head.trim = 0;
  head.apply_sigmoid = 0;
  EXA_LAUNCH(CAT_MISC_FWD, launch_head(u4_, head, s));  // outc, unet3d.py:318
  forward_valid_ = true;
  return Status::OK();
}

// BatchNorm + LeakyReLU backward, weight / bias gradients, data gradient of one conv layer
Status Trainer::layer_backward(int l, const TView& grad_a, const float* x, float* grads,
                               cudaStream_t s) {
  Layer& L = layers_[l];
  const double count = (double)L.z.voxels();
  Act dz = dz_;
  dz.D = L.z.D; dz.H = L.z.H; dz.W = L.z.W;
  dz.C = dz.cstride = L.cout;
  EXA_CUDA(cudaMemsetAsync(L.sums, 0, sizeof(double) * 2 * L.cout, s));
  EXA_LAUNCH(CAT_BN_BWD, launch_bn_bwd_reduce(grad_a, L.scale, L.shift, encoded(L.z), L.mean, L.rstd, L.sums, s));
  EXA_LAUNCH(CAT_BN_BWD, launch_bn_bwd_finalize(L.sums, L.cout, count, L.gamma, L.rstd, grads + L.ggamma,
                                    grads + L.gbeta, L.coef, s));
  EXA_CUDA(cudaMemsetAsync(L.sums, 0, sizeof(double) * 2 * L.cout, s));
  EXA_LAUNCH(CAT_BN_BWD, launch_bn_bwd_apply(grad_a, L.scale, L.shift, encoded(L.z), L.mean, L.rstd, L.coef, dz, L.sums, s));
  EXA_LAUNCH(CAT_BN_BWD, launch_double_to_float(L.sums, grads + L.gb, L.cout, s));
  if (l == 0) {
    EXA_LAUNCH(CAT_WGRAD_STEM, launch_wgrad_stem(x, dz, partial_, num_sms_, s));
    EXA_LAUNCH(CAT_WGRAD_REDUCE, launch_wgrad_reduce(partial_, wgrad_stem_splits(dz, num_sms_), (size_t)L.cout * 27,
                                   grads + L.gw, s));
    return Status::OK();
  }
  EXA_LAUNCH(CAT_WGRAD, launch_wgrad(L.x, dz, partial_, num_sms_, s));
  EXA_LAUNCH(CAT_WGRAD_REDUCE, launch_wgrad_reduce(partial_, wgrad_splits(L.x, L.cout, num_sms_),
                                 (size_t)L.cout * L.cin * 27, grads + L.gw, s));
  EXA_LAUNCH(CAT_DGRAD, conv_any(dz, L.gx, L.w_bwd, L.w_bwd_zf, zero_bias_, s));
  return Status::OK();
}

Status Trainer::backward(const float* x, const float* dlogits, float* grads, cudaStream_t s) {
  EXA_CUDA(cudaSetDevice(device_));
  EXA_CHECK(x && dlogits && grads, "train_backward: null argument");
  EXA_CHECK(forward_valid_, "train_backward: no forward pass to differentiate (call "
                            "exa_train_forward first; parameters must not be re-bound in between)");
  const int* c = chan_;
  const int C = out_channels_;
  // head (unet3d.py:318)
  const int nh = C * c[0] + C;
  EXA_CUDA(cudaMemsetAsync(head_sums_, 0, sizeof(double) * nh, s));
  EXA_LAUNCH(CAT_HEAD_BWD, launch_head_bwd_dw(dlogits, u4_, C, head_sums_, s));
  EXA_LAUNCH(CAT_HEAD_BWD, launch_double_to_float(head_sums_, grads + g_head_w_, C * c[0], s));
  EXA_LAUNCH(CAT_HEAD_BWD, launch_double_to_float(head_sums_ + C * c[0], grads + g_head_b_, C, s));
  EXA_LAUNCH(CAT_HEAD_BWD, launch_head_bwd_dx(dlogits, head_w_, C, g_u4_, s));
  // decoder
  EXA_TRY(layer_backward(17, plain(g_u4_), x, grads, s));
  EXA_TRY(layer_backward(16, encoded(g_u4a_), x, grads, s));
  EXA_LAUNCH(CAT_UPSAMPLE_BWD, launch_upsample_bwd(encoded(channel_slice(g_cat4_, c[0], c[0])), g_u3_, s));
  EXA_TRY(layer_backward(15, plain(g_u3_), x, grads, s));
  EXA_TRY(layer_backward(14, encoded(g_u3a_), x, grads, s));
  EXA_LAUNCH(CAT_UPSAMPLE_BWD, launch_upsample_bwd(encoded(channel_slice(g_cat3_, c[1], c[1])), g_u2_, s));
  EXA_TRY(layer_backward(13, plain(g_u2_), x, grads, s));
  EXA_TRY(layer_backward(12, encoded(g_u2a_), x, grads, s));
  EXA_LAUNCH(CAT_UPSAMPLE_BWD, launch_upsample_bwd(encoded(channel_slice(g_cat2_, c[2], c[2])), g_u1_, s));
  EXA_TRY(layer_backward(11, plain(g_u1_), x, grads, s));
  EXA_TRY(layer_backward(10, encoded(g_u1a_), x, grads, s));
  EXA_LAUNCH(CAT_UPSAMPLE_BWD, launch_upsample_bwd(encoded(channel_slice(g_cat1_, c[3], c[3])), g_x5_, s));
  // encoder: every skip tensor collects its concat half and the max-pool's gradient
  EXA_TRY(layer_backward(9, plain(g_x5_), x, grads, s));
  EXA_TRY(layer_backward(8, encoded(g_d4a_), x, grads, s));
  EXA_LAUNCH(CAT_POOL_BWD, launch_pool_bwd_merge(encoded(channel_slice(g_cat1_, 0, c[3])), encoded(g_p4_), x4_,
                                   g_x4_, s));
  EXA_TRY(layer_backward(7, plain(g_x4_), x, grads, s));
  EXA_TRY(layer_backward(6, encoded(g_d3a_), x, grads, s));
  EXA_LAUNCH(CAT_POOL_BWD, launch_pool_bwd_merge(encoded(channel_slice(g_cat2_, 0, c[2])), encoded(g_p3_), x3_,
                                   g_x3_, s));
  EXA_TRY(layer_backward(5, plain(g_x3_), x, grads, s));
  EXA_TRY(layer_backward(4, encoded(g_d2a_), x, grads, s));
  EXA_LAUNCH(CAT_POOL_BWD, launch_pool_bwd_merge(encoded(channel_slice(g_cat3_, 0, c[1])), encoded(g_p2_), x2_,
                                   g_x2_, s));
  EXA_TRY(layer_backward(3, plain(g_x2_), x, grads, s));
  EXA_TRY(layer_backward(2, encoded(g_d1a_), x, grads, s));
  EXA_LAUNCH(CAT_POOL_BWD, launch_pool_bwd_merge(encoded(channel_slice(g_cat4_, 0, c[0])), encoded(g_p1_), x1_,
                                   g_x1_, s));
  EXA_TRY(layer_backward(1, plain(g_x1_), x, grads, s));
  EXA_TRY(layer_backward(0, encoded(g_a0_), x, grads, s));
  return Status::OK();
}

// ---------------------------------------------------------------------------
// operator-level entry points (tests)
// ---------------------------------------------------------------------------
namespace {
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { free_dev(p); }
  Status alloc(size_t bytes) {
    EXA_CUDA(cudaMalloc(&p, bytes ? bytes : 16));
    return Status::OK();
  }
};
Status device_sms(int device, int* sms) {
  int count = 0;
  EXA_CHECK(cudaGetDeviceCount(&count) == cudaSuccess && device >= 0 && device < count,
            "no such CUDA device (this library has no CPU fallback)");
  EXA_CUDA(cudaSetDevice(device));
  EXA_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, device));
  return Status::OK();
}
Act dense_act(const void* p, int B, int D, int H, int W, int C, bool f32) {
  Act a;
  a.ptr = const_cast<void*>(p);
  a.B = B; a.D = D; a.H = H; a.W = W;
  a.C = a.cstride = C;
  a.coff = 0;
  a.fp32 = f32;
  return a;
}
}  // namespace

Status conv3d_weight_grad(int device, int precision, const void* x, const void* dz, int B, int D,
                          int H, int W, int cin, int cout, float* dw, cudaStream_t s) {
  EXA_CHECK(x && dz && dw && B > 0 && D > 0 && H > 0 && W > 0, "conv3d_weight_grad: bad arguments");
  EXA_CHECK(precision == EXA_PRECISION_BF16 || precision == EXA_PRECISION_FP32, "unknown precision");
  int sms = 0;
  EXA_TRY(device_sms(device, &sms));
  const bool f32 = precision == EXA_PRECISION_FP32;
  const Act dza = dense_act(dz, B, D, H, W, cout, f32);
  DevBuf partial;
  if (cin == 1) {  // stem: x is the raw float32 (B,1,D,H,W) input
    const int splits = wgrad_stem_splits(dza, sms);
    EXA_TRY(partial.alloc((size_t)splits * cout * 27 * 4));
    EXA_TRY(launch_wgrad_stem((const float*)x, dza, (float*)partial.p, sms, s));
    EXA_TRY(launch_wgrad_reduce((const float*)partial.p, splits, (size_t)cout * 27, dw, s));
  } else {
    const Act xa = dense_act(x, B, D, H, W, cin, f32);
    EXA_TRY(partial.alloc(wgrad_partial_elems(xa, cout, sms) * 4));
    EXA_TRY(launch_wgrad(xa, dza, (float*)partial.p, sms, s));
    EXA_TRY(launch_wgrad_reduce((const float*)partial.p, wgrad_splits(xa, cout, sms),
                                (size_t)cout * cin * 27, dw, s));
  }
  EXA_CUDA(cudaStreamSynchronize(s));
  return Status::OK();
}

Status conv3d_data_grad(int device, int precision, const void* dz, const float* w, int B, int D,
                        int H, int W, int cin, int cout, void* dx, cudaStream_t s) {
  EXA_CHECK(dz && w && dx && B > 0 && D > 0 && H > 0 && W > 0, "conv3d_data_grad: bad arguments");
  EXA_CHECK(precision == EXA_PRECISION_BF16 || precision == EXA_PRECISION_FP32, "unknown precision");
  EXA_CHECK(cin % 32 == 0 && cout % 32 == 0, "conv3d_data_grad: channels must be multiples of 32");
  int sms = 0;
  EXA_TRY(device_sms(device, &sms));
  const bool f32 = precision == EXA_PRECISION_FP32;
  const size_t esz = f32 ? 4 : 2, n = (size_t)27 * cin * cout;
  const Act dza = dense_act(dz, B, D, H, W, cout, f32);
  DevBuf wp, wz, zero, raw;
  EXA_TRY(wp.alloc(n * esz));
  EXA_TRY(zero.alloc(4 * (size_t)cin));
  EXA_TRY(raw.alloc(dza.voxels() * cin * esz));
  EXA_CUDA(cudaMemsetAsync(zero.p, 0, 4 * (size_t)cin, s));
  EXA_TRY(launch_pack_conv_weights(w, wp.p, cout, cin, f32, true, false, f32, s));
  const Act rawa = dense_act(raw.p, B, D, H, W, cin, f32);
  static const bool no_zf = getenv("EXA_TRAIN_NO_ZFOLD") != nullptr;
  if (f32) {
    EXA_TRY(launch_conv_fp32(dza, rawa, (const float*)wp.p, (const float*)zero.p, s));
  } else if (!no_zf && conv_zfold_supported(dza, cin, true)) {
    EXA_TRY(wz.alloc(n * esz));
    EXA_TRY(launch_pack_conv_weights(w, wz.p, cout, cin, false, true, true, false, s));
    EXA_TRY(launch_conv_zfold(dza, rawa, (const __nv_bfloat16*)wz.p, (const float*)zero.p, nullptr,
                              nullptr, nullptr, sms, true, s));
  } else {
    EXA_TRY(launch_conv_umma(dza, rawa, (const __nv_bfloat16*)wp.p, (const float*)zero.p, nullptr,
                             sms, s));
  }
  TView enc;
  enc.a = rawa;
  enc.enc = true;
  EXA_TRY(launch_decode(enc, dense_act(dx, B, D, H, W, cin, f32), s));
  EXA_CUDA(cudaStreamSynchronize(s));
  return Status::OK();
}

}  // namespace exa
