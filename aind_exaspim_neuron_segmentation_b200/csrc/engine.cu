// Engine implementation: weight folding/packing, workspace layout, the layer schedule of
// reference unet3d.py:77-105 and the volume driver of reference inference.py:29-126.
#include "engine.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

namespace exa {

// ---------------------------------------------------------------------------
// pure host helpers
// ---------------------------------------------------------------------------
// range(0, dim - patch + stride, stride)                       (inference.py:389-393)
std::vector<int> axis_starts(int dim, int patch, int overlap) {
  std::vector<int> out;
  const int stride = patch - overlap;
  if (stride <= 0) return out;
  for (int s = 0; s < dim - patch + stride; s += stride) out.push_back(s);
  return out;
}

static Status check_params(const exa_predict_params& p) {
  for (int i = 0; i < 3; ++i) {
    EXA_CHECK(p.patch[i] > 0 && p.patch[i] % 16 == 0,
              "patch_shape entries must be positive multiples of 16 (unet3d.py:281-288)");
    EXA_CHECK(p.overlap[i] >= 0 && p.overlap[i] < p.patch[i], "overlap must be in [0, patch)");
    EXA_CHECK(p.patch[i] - 2 * p.trim > 0, "trim too large for patch_shape");
  }
  EXA_CHECK(p.trim >= 0, "trim must be >= 0");
  EXA_CHECK(p.brightness_clip >= 0 && p.brightness_clip <= 65535,
            "brightness_clip must be in [0, 65535]");
  EXA_CHECK(p.pct_lo >= 0 && p.pct_lo <= 100 && p.pct_hi >= 0 && p.pct_hi <= 100,
            "percentiles must be in [0, 100]");
  return Status::OK();
}

Status make_plan(int D, int H, int W, const exa_predict_params& p, Plan* plan) {
  EXA_TRY(check_params(p));
  EXA_CHECK(D > 0 && H > 0 && W > 0, "volume dims must be positive");
  plan->D = D;
  plan->H = H;
  plan->W = W;
  const int dims[3] = {D, H, W};
  AxisGeom* g[3] = {&plan->az, &plan->ay, &plan->ax};
  for (int i = 0; i < 3; ++i) {
    g[i]->dim = dims[i];
    g[i]->patch = p.patch[i];
    g[i]->stride = p.patch[i] - p.overlap[i];
    g[i]->trim = p.trim;
    g[i]->n = (int)axis_starts(dims[i], p.patch[i], p.overlap[i]).size();
  }
  plan->n_patches = plan->az.n * plan->ay.n * plan->ax.n;
  return Status::OK();
}

Status plan_slab(const Plan& plan, int row_begin, int row_end, exa_slab_plan* out) {
  const AxisGeom& z = plan.az;
  EXA_CHECK(row_begin >= 0 && row_begin <= row_end && row_end <= z.n, "row range out of bounds");
  const int keep = z.patch - 2 * z.trim;
  memset(out, 0, sizeof(*out));
  out->nz = z.n;
  out->ny = plan.ay.n;
  out->nx = plan.ax.n;
  out->n_patches = plan.n_patches;
  if (row_begin == row_end) return Status::OK();  // empty slab: all ranges empty
  // a plane may be covered by at most two consecutive z rows, otherwise more than the
  // neighbouring slab would have to contribute partial sums
  EXA_CHECK(keep <= 2 * z.stride || (row_begin == 0 && row_end == z.n),
            "slab sharding needs patch - 2*trim <= 2*(patch - overlap) along z");
  out->in_z0 = z.stride * row_begin;
  out->in_z1 = std::min(z.stride * (row_end - 1) + z.patch, z.dim);
  out->out_z0 = row_begin == 0 ? 0 : z.stride * row_begin + z.trim;
  out->out_z1 = row_end == z.n ? z.dim : std::min(z.stride * row_end + z.trim, z.dim);
  if (row_end < z.n) {
    out->halo_z0 = std::min(z.stride * row_end + z.trim, z.dim);
    out->halo_z1 = std::max(out->halo_z0, std::min(z.stride * (row_end - 1) + z.trim + keep, z.dim));
  }
  if (row_begin > 0) {
    out->seed_z0 = std::min(z.stride * row_begin + z.trim, z.dim);
    out->seed_z1 =
        std::max(out->seed_z0, std::min(z.stride * (row_begin - 1) + z.trim + keep, z.dim));
  }
  return Status::OK();
}

// np.percentile(a, q) with the default method="linear" on the sorted multiset described by
// `hist` (img_util.py:526).  Mirrors numpy (>= 1.22, checked against 2.3) step by step so
// that the float64 results are bit-identical: q = pct/100, virtual index (n-1)*q,
// gamma = index - floor(index), neighbours clamped as in numpy's _get_indexes, and numpy's
// two-sided _lerp.
static double percentile_linear(const uint64_t* hist, int nbins, uint64_t n, double pct) {
  const double q = pct / 100.0;
  const double vidx = (double)(n - 1) * q;
  const double prev = floor(vidx);
  const double gamma = vidx - prev;
  const int64_t last = (int64_t)n - 1;
  int64_t i0 = (int64_t)prev;
  int64_t i1 = i0 + 1;
  if (vidx >= (double)last) i0 = i1 = last;
  if (vidx < 0) i0 = i1 = 0;
  // order statistics i0, i1 from the histogram
  double a = 0, b = 0;
  uint64_t cum = 0;
  bool got_a = false, got_b = false;
  for (int v = 0; v < nbins && !(got_a && got_b); ++v) {
    cum += hist[v];
    if (!got_a && (uint64_t)i0 < cum) {
      a = (double)v;
      got_a = true;
    }
    if (!got_b && (uint64_t)i1 < cum) {
      b = (double)v;
      got_b = true;
    }
  }
  const double diff = b - a;
  double r = a + diff * gamma;
  if (gamma >= 0.5) r = b - diff * (1.0 - gamma);
  return r;
}

Status percentiles_from_hist(const uint64_t* hist, int nbins, double q_lo, double q_hi, double* mn,
                             double* mx) {
  EXA_CHECK(hist && nbins > 0 && mn && mx, "percentiles_from_hist: bad arguments");
  uint64_t n = 0;
  for (int i = 0; i < nbins; ++i) n += hist[i];
  EXA_CHECK(n > 0, "percentiles_from_hist: empty histogram");
  *mn = percentile_linear(hist, nbins, n, q_lo);
  *mx = percentile_linear(hist, nbins, n, q_hi);
  return Status::OK();
}

// ---------------------------------------------------------------------------
// Engine: lifetime
// ---------------------------------------------------------------------------
static void free_dev(void* p) {
  if (p) cudaFree(p);
}

Engine::~Engine() {
  cudaSetDevice(device_);
  for (auto& r : prof_) {
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  for (auto& L : layers_) {
    free_dev(L.w_bf16);
    free_dev(L.w_f32);
    free_dev(L.w_zfold);
    free_dev(L.bias);
  }
  for (auto& u : upconv_) {
    free_dev(u.w);
    free_dev(u.bias);
  }
  free_dev(head_w_);
  free_dev(head_b_);
  free_dev(stem_band_);
  free_dev(stem_bias_);
  free_dev(ws_);
  free_dev(lut_);
  free_dev(probs_);
  free_dev(starts_dev_);
  free_dev(hist_dev_);
  free_dev(seed_);
  free_dev(probs_hold_);
  free_dev(probs_alt_);
  if (stitch_stream_) {
    cudaStreamDestroy(stitch_stream_);
    cudaEventDestroy(conv_done_);
    cudaEventDestroy(probs_free_);
    cudaEventDestroy(alt_free_);
  }
  if (copy_event_) cudaEventDestroy(copy_event_);
  if (peer_ready_) {
    cudaEventDestroy(peer_ready_);
    for (int i = 0; i < kPeerStreams; ++i) {
      cudaStreamDestroy(peer_stream_[i]);
      cudaEventDestroy(peer_done_[i]);
    }
  }
  free_dev(vol_stage_);
  free_dev(out_stage_);
  if (copy_stream_) cudaStreamDestroy(copy_stream_);
}

Status Engine::init() {
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
  const char* nz = getenv("EXA_NO_ZFOLD");
  use_zfold_ = !(nz && nz[0] == '1');
  const char* ns = getenv("EXA_NO_TC_STEM");
  use_tc_stem_ = !(ns && ns[0] == '1');
  const char* so = getenv("EXA_STITCH_OVERLAP");
  overlap_stitch_ = so && so[0] == '1';
  const char* np = getenv("EXA_NO_PAIR");
  use_pair_ = !(np && np[0] == '1');
  return Status::OK();
}

// ---------------------------------------------------------------------------
// Engine: profiling scopes
// ---------------------------------------------------------------------------
Engine::Scope::Scope(Engine* eng, int cat, cudaStream_t st) : e(eng), s(st) {
  ++e->launches;
  if (!e->prof_on_) return;
  ProfRec r;
  r.cat = cat;
  r.tag = cat == CAT_CONV ? e->cur_tag_ : -1;
  if (cudaEventCreate(&r.start) != cudaSuccess || cudaEventCreate(&r.stop) != cudaSuccess) return;
  cudaEventRecord(r.start, s);
  stop = r.stop;
  e->prof_.push_back(r);
}
Engine::Scope::~Scope() {
  if (stop) cudaEventRecord(stop, s);
}

Status Engine::profile_begin() {
  EXA_CUDA(cudaSetDevice(device_));
  for (auto& r : prof_) {
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  prof_.clear();
  prof_on_ = true;
  return Status::OK();
}

Status Engine::profile_end(double* ms_by_cat, int64_t* launches_by_cat, int n) {
  EXA_CUDA(cudaSetDevice(device_));
  EXA_CHECK(ms_by_cat && launches_by_cat && n >= CAT_COUNT, "profile_end: need 7 categories");
  prof_on_ = false;
  for (int i = 0; i < n; ++i) {
    ms_by_cat[i] = 0;
    launches_by_cat[i] = 0;
  }
  Status st = Status::OK();
  double* layer_ms = layer_ms_;
  int64_t* layer_n = layer_n_;
  for (int i = 0; i < 18; ++i) {
    layer_ms[i] = 0;
    layer_n[i] = 0;
  }
  for (auto& r : prof_) {
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(r.stop);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.start, r.stop);
    if (e != cudaSuccess && st.ok) st = Status::Err(std::string("cudaEventElapsedTime: ") + cudaGetErrorString(e));
    ms_by_cat[r.cat] += ms;
    launches_by_cat[r.cat] += 1;
    if (r.tag >= 0 && r.tag < 18) {
      layer_ms[r.tag] += ms;
      layer_n[r.tag] += 1;
    }
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  prof_.clear();
  if (getenv("EXA_LAYER_PROF")) {
    for (int i = 0; i < 18; ++i)
      if (layer_n[i])
        fprintf(stderr, "[layer %2d %-36s %3d->%3d] %4d launches, %.3f ms total, %.4f ms/launch\n", i,
                layers_[i].conv_key.c_str(), layers_[i].cin, layers_[i].cout, (int)layer_n[i],
                layer_ms[i], layer_ms[i] / layer_n[i]);
  }
  return st;
}

Status Engine::profile_layers(double* ms, int64_t* launches, int32_t* kind, int n) const {
  EXA_CHECK(ms && launches && kind && n >= 18, "profile_layers: need 18 entries");
  for (int i = 0; i < 18; ++i) {
    ms[i] = layer_ms_[i];
    launches[i] = layer_n_[i];
    kind[i] = layer_kind_[i];
  }
  return Status::OK();
}

// ---------------------------------------------------------------------------
// Engine: weights (replaces UNet3D.load_state_dict, inference.py:420-421)
// ---------------------------------------------------------------------------
Status Engine::load_weight(const char* name, const void* data, const int64_t* shape, int ndim,
                           int dtype) {
  EXA_CHECK(name && (data || ndim == 0 || true), "load_weight: null argument");
  EXA_CHECK(ndim >= 0 && ndim <= 5, "load_weight: ndim must be <= 5");
  EXA_CHECK(dtype == EXA_DTYPE_F32 || dtype == EXA_DTYPE_I64, "load_weight: unknown dtype");
  HostTensor t;
  t.dtype = dtype;
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    EXA_CHECK(shape[i] >= 0, "load_weight: negative dim");
    t.shape.push_back(shape[i]);
    n *= shape[i];
  }
  EXA_CHECK(data != nullptr || n == 0, "load_weight: null data");
  if (dtype == EXA_DTYPE_F32) {
    t.f32.assign((const float*)data, (const float*)data + n);
  } else {
    t.i64.assign((const int64_t*)data, (const int64_t*)data + n);
  }
  raw_[name] = std::move(t);
  finalized_ = false;
  return Status::OK();
}

namespace {

struct LayerSpec {
  const char* prefix;
  int conv_idx, bn_idx;
  int cin, cout;
};
// unet3d.py:56-74: channels c[k] = int(32 * 2^k * width_multiplier); trilinear upsampling halves the
// deep widths (factor 2) and gives the decoder's DoubleConv a mid width of in/2 (unet3d.py:248-258).
// width_multiplier = 1, trilinear = True is what load_model builds (inference.py:419-420).
void layer_table(const int c[5], bool trilinear, LayerSpec out[18]) {
  const int f = trilinear ? 2 : 1;
  static const char* kPrefix[9] = {
      "inc.double_conv",
      "down1.maxpool_conv.1.double_conv", "down2.maxpool_conv.1.double_conv",
      "down3.maxpool_conv.1.double_conv", "down4.maxpool_conv.1.double_conv",
      "up1.conv.double_conv", "up2.conv.double_conv", "up3.conv.double_conv", "up4.conv.double_conv"};
  // (in, mid, out) of the nine DoubleConv blocks
  int io[9][3] = {{1, c[0], c[0]},
                  {c[0], c[1], c[1]}, {c[1], c[2], c[2]}, {c[2], c[3], c[3]},
                  {c[3], c[4] / f, c[4] / f},
                  {c[4], trilinear ? c[4] / 2 : c[3] / f, c[3] / f},
                  {c[3], trilinear ? c[3] / 2 : c[2] / f, c[2] / f},
                  {c[2], trilinear ? c[2] / 2 : c[1] / f, c[1] / f},
                  {c[1], trilinear ? c[1] / 2 : c[0], c[0]}};
  for (int b = 0; b < 9; ++b) {
    out[2 * b] = LayerSpec{kPrefix[b], 0, 1, io[b][0], io[b][1]};
    out[2 * b + 1] = LayerSpec{kPrefix[b], 3, 4, io[b][1], io[b][2]};
  }
}

uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)((u >> 16) | ((u & 0xFFFFu) ? 0x40 : 0));
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

}  // namespace

Status Engine::finalize_weights() {
  EXA_CUDA(cudaSetDevice(device_));
  std::map<std::string, bool> used;
  auto get = [&](const std::string& key, std::vector<int64_t> shape, int dtype,
                 const HostTensor** out) -> Status {
    auto it = raw_.find(key);
    EXA_CHECK(it != raw_.end(), "missing state_dict entry: " + key);
    const HostTensor& t = it->second;
    EXA_CHECK(t.dtype == dtype, "wrong dtype for state_dict entry: " + key);
    EXA_CHECK(t.shape == shape, "wrong shape for state_dict entry: " + key);
    used[key] = true;
    *out = &t;
    return Status::OK();
  };

  // configuration from the entries themselves: width of the first conv (32 * width_multiplier) and
  // the presence of transposed-conv weights (trilinear=False, unet3d.py:254-256)
  {
    auto it = raw_.find("inc.double_conv.0.weight");
    EXA_CHECK(it != raw_.end(), "missing state_dict entry: inc.double_conv.0.weight");
    EXA_CHECK(it->second.shape.size() == 5 && it->second.shape[0] >= 32 &&
                  it->second.shape[0] <= 128 && it->second.shape[0] % 32 == 0,
              "unsupported width: inc.double_conv.0.weight must have 32, 64, 96 or 128 output channels "
              "(width_multiplier 1..4)");
    for (int k = 0; k < 5; ++k) chan_[k] = (int)it->second.shape[0] << k;
    trilinear_ = raw_.find("up1.up.weight") == raw_.end();
  }
  LayerSpec specs[18];
  layer_table(chan_, trilinear_, specs);
  const int c0 = chan_[0];
  // head: its shape decides affinity (3) vs foreground (1) mode  (unet3d.py:75,318)
  {
    auto it = raw_.find("outc.conv.weight");
    EXA_CHECK(it != raw_.end(), "missing state_dict entry: outc.conv.weight");
    EXA_CHECK(it->second.shape.size() == 5 && it->second.shape[1] == c0 &&
                  it->second.shape[2] == 1 && it->second.shape[3] == 1 &&
                  it->second.shape[4] == 1 && it->second.shape[0] >= 1 &&
                  it->second.shape[0] <= 8,
              "wrong shape for state_dict entry: outc.conv.weight");
    out_channels_ = (int)it->second.shape[0];
  }
  const HostTensor *hw = nullptr, *hb = nullptr;
  EXA_TRY(get("outc.conv.weight", {out_channels_, c0, 1, 1, 1}, EXA_DTYPE_F32, &hw));
  EXA_TRY(get("outc.conv.bias", {out_channels_}, EXA_DTYPE_F32, &hb));
  free_dev(head_w_);
  free_dev(head_b_);
  head_w_ = head_b_ = nullptr;
  head_w_host_ = hw->f32;
  head_b_host_ = hb->f32;
  EXA_CUDA(cudaMalloc(&head_w_, sizeof(float) * out_channels_ * c0));
  EXA_CUDA(cudaMalloc(&head_b_, sizeof(float) * out_channels_));
  EXA_CUDA(cudaMemcpy(head_w_, hw->f32.data(), sizeof(float) * out_channels_ * c0,
                      cudaMemcpyHostToDevice));
  EXA_CUDA(cudaMemcpy(head_b_, hb->f32.data(), sizeof(float) * out_channels_,
                      cudaMemcpyHostToDevice));

  // transposed convs of the decoder (trilinear=False): ConvTranspose3d(in, in/2, k=2, s=2), weight
  // (in, in/2, 2, 2, 2) -> [tap = dz*4+dy*2+dx][cin][cout], bf16-rounded once in bf16 mode
  for (int u = 0; u < 4; ++u) {
    free_dev(upconv_[u].w);
    free_dev(upconv_[u].bias);
    upconv_[u] = UpConv{};
    if (trilinear_) continue;
    const int cin = chan_[4 - u], cout = cin / 2;
    const std::string key = "up" + std::to_string(u + 1) + ".up";
    const HostTensor *w, *b;
    EXA_TRY(get(key + ".weight", {cin, cout, 2, 2, 2}, EXA_DTYPE_F32, &w));
    EXA_TRY(get(key + ".bias", {cout}, EXA_DTYPE_F32, &b));
    upconv_[u].cin = cin;
    upconv_[u].cout = cout;
    const size_t n = (size_t)8 * cin * cout;
    std::vector<float> pk(n);
    for (int ci = 0; ci < cin; ++ci)
      for (int co = 0; co < cout; ++co)
        for (int tap = 0; tap < 8; ++tap) {
          float v = w->f32[((size_t)ci * cout + co) * 8 + tap];
          if (precision_ == EXA_PRECISION_BF16) {
            const uint32_t bits = (uint32_t)f32_to_bf16_rn(v) << 16;
            memcpy(&v, &bits, 4);
          }
          pk[((size_t)tap * cin + ci) * cout + co] = v;
        }
    EXA_CUDA(cudaMalloc(&upconv_[u].w, n * 4));
    EXA_CUDA(cudaMemcpy(upconv_[u].w, pk.data(), n * 4, cudaMemcpyHostToDevice));
    EXA_CUDA(cudaMalloc(&upconv_[u].bias, sizeof(float) * cout));
    EXA_CUDA(cudaMemcpy(upconv_[u].bias, b->f32.data(), sizeof(float) * cout, cudaMemcpyHostToDevice));
  }

  for (int li = 0; li < 18; ++li) {
    const LayerSpec& sp = specs[li];
    ConvLayer& L = layers_[li];
    L.conv_key = std::string(sp.prefix) + "." + std::to_string(sp.conv_idx);
    L.bn_key = std::string(sp.prefix) + "." + std::to_string(sp.bn_idx);
    L.cin = sp.cin;
    L.cout = sp.cout;
    const HostTensor *w, *b, *g, *be, *mu, *var, *nbt;
    EXA_TRY(get(L.conv_key + ".weight", {sp.cout, sp.cin, 3, 3, 3}, EXA_DTYPE_F32, &w));
    EXA_TRY(get(L.conv_key + ".bias", {sp.cout}, EXA_DTYPE_F32, &b));
    EXA_TRY(get(L.bn_key + ".weight", {sp.cout}, EXA_DTYPE_F32, &g));
    EXA_TRY(get(L.bn_key + ".bias", {sp.cout}, EXA_DTYPE_F32, &be));
    EXA_TRY(get(L.bn_key + ".running_mean", {sp.cout}, EXA_DTYPE_F32, &mu));
    EXA_TRY(get(L.bn_key + ".running_var", {sp.cout}, EXA_DTYPE_F32, &var));
    EXA_TRY(get(L.bn_key + ".num_batches_tracked", {}, EXA_DTYPE_I64, &nbt));

    // eval-mode BatchNorm3d folded into the conv: y = (conv(x)+b-mu)*g/sqrt(var+eps)+beta
    // (unet3d.py:144,147; eps = 1e-5 default)
    std::vector<double> scale(sp.cout);
    std::vector<float> bias(sp.cout);
    for (int co = 0; co < sp.cout; ++co) {
      scale[co] = (double)g->f32[co] / sqrt((double)var->f32[co] + 1e-5);
      bias[co] = (float)(((double)b->f32[co] - (double)mu->f32[co]) * scale[co] +
                         (double)be->f32[co]);
    }
    free_dev(L.bias);
    free_dev(L.w_bf16);
    free_dev(L.w_f32);
    free_dev(L.w_zfold);
    L.w_zfold = nullptr;
    L.bias = nullptr;
    L.w_bf16 = nullptr;
    L.w_f32 = nullptr;
    EXA_CUDA(cudaMalloc(&L.bias, sizeof(float) * sp.cout));
    EXA_CUDA(cudaMemcpy(L.bias, bias.data(), sizeof(float) * sp.cout, cudaMemcpyHostToDevice));

    auto wsrc = [&](int co, int ci, int tap) -> double {
      return (double)w->f32[((size_t)co * sp.cin + ci) * 27 + tap] * scale[co];
    };
    if (li == 0) {
      // the stem kernels produce 32 channels per pass: one weight group per 32 output channels
      const int groups = sp.cout / 32;
      stem_.assign(groups, StemWeights{});
      for (int g = 0; g < groups; ++g) {
        for (int tap = 0; tap < 27; ++tap)
          for (int co = 0; co < 32; ++co) stem_[g].w[tap][co] = (float)wsrc(g * 32 + co, 0, tap);
        for (int co = 0; co < 32; ++co) stem_[g].b[co] = bias[g * 32 + co];
      }
      if (precision_ == EXA_PRECISION_BF16) {
        // Toeplitz (band) form for the tensor-core stem (conv_stem.cuh): B[kz,ky][n = xo*32 + c]
        // [k = 2*x' + part] = w[c][kz][ky][kx = x' - xo] for both the hi and the lo part of the
        // input, bf16-rounded once like every other layer's weights
        const size_t per_group = (size_t)9 * 128 * 16;
        std::vector<uint16_t> band(per_group * groups, 0);
        for (int g = 0; g < groups; ++g)
          for (int t9 = 0; t9 < 9; ++t9)
            for (int xo = 0; xo < 4; ++xo)
              for (int c = 0; c < 32; ++c)
                for (int kx = 0; kx < 3; ++kx)
                  for (int part = 0; part < 2; ++part)
                    band[g * per_group + ((size_t)t9 * 128 + xo * 32 + c) * 16 + 2 * (xo + kx) + part] =
                        f32_to_bf16_rn(stem_[g].w[t9 * 3 + kx][c]);
        free_dev(stem_band_);
        free_dev(stem_bias_);
        stem_band_ = nullptr;
        stem_bias_ = nullptr;
        EXA_CUDA(cudaMalloc(&stem_band_, band.size() * 2));
        EXA_CUDA(cudaMemcpy(stem_band_, band.data(), band.size() * 2, cudaMemcpyHostToDevice));
        EXA_CUDA(cudaMalloc(&stem_bias_, sizeof(float) * sp.cout));
        EXA_CUDA(cudaMemcpy(stem_bias_, bias.data(), sizeof(float) * sp.cout, cudaMemcpyHostToDevice));
      }
      continue;
    }
    const size_t n = (size_t)27 * sp.cin * sp.cout;
    if (precision_ == EXA_PRECISION_BF16) {
      std::vector<uint16_t> pk(n);  // [tap][cout][cin]: K-major B operand of the implicit GEMM
      for (int tap = 0; tap < 27; ++tap)
        for (int co = 0; co < sp.cout; ++co)
          for (int ci = 0; ci < sp.cin; ++ci)
            pk[((size_t)tap * sp.cout + co) * sp.cin + ci] = f32_to_bf16_rn((float)wsrc(co, ci, tap));
      EXA_CUDA(cudaMalloc(&L.w_bf16, n * 2));
      EXA_CUDA(cudaMemcpy(L.w_bf16, pk.data(), n * 2, cudaMemcpyHostToDevice));
      if ((sp.cin == 32 || sp.cin == 64 || sp.cin == 128) && (sp.cout == 32 || sp.cout == 64)) {
        // z-folded layout: B operand rows of one in-plane tap are [kz=2 | kz=1 | kz=0] x cout
        std::vector<uint16_t> zf(n);
        for (int t9 = 0; t9 < 9; ++t9)
          for (int kzr = 0; kzr < 3; ++kzr)
            for (int co = 0; co < sp.cout; ++co)
              for (int ci = 0; ci < sp.cin; ++ci)
                zf[(((size_t)t9 * 3 + kzr) * sp.cout + co) * sp.cin + ci] =
                    pk[((size_t)((2 - kzr) * 9 + t9) * sp.cout + co) * sp.cin + ci];
        EXA_CUDA(cudaMalloc(&L.w_zfold, n * 2));
        EXA_CUDA(cudaMemcpy(L.w_zfold, zf.data(), n * 2, cudaMemcpyHostToDevice));
      }
    } else {
      std::vector<float> pk(n);  // [tap][cin][cout]
      for (int tap = 0; tap < 27; ++tap)
        for (int ci = 0; ci < sp.cin; ++ci)
          for (int co = 0; co < sp.cout; ++co)
            pk[((size_t)tap * sp.cin + ci) * sp.cout + co] = (float)wsrc(co, ci, tap);
      EXA_CUDA(cudaMalloc(&L.w_f32, n * 4));
      EXA_CUDA(cudaMemcpy(L.w_f32, pk.data(), n * 4, cudaMemcpyHostToDevice));
    }
  }
  // strict: no unexpected keys (load_state_dict(strict=True), inference.py:421)
  for (auto& kv : raw_) {
    EXA_CHECK(used.count(kv.first), "unexpected state_dict entry: " + kv.first);
  }
  finalized_ = true;
  return Status::OK();
}

// ---------------------------------------------------------------------------
// Engine: workspace
// ---------------------------------------------------------------------------
namespace {
struct WsLayout {
  // element offsets (in units of the activation element) of every buffer
  size_t xhi, xlo, a0, cat4, p1, d1a, cat3, p2, d2a, cat2, p3, d3a, cat1, p4, d4a, x5, u1a, u1, u2a, u2, u3a,
      u3, total;
};
WsLayout layout(int B, int pz, int py, int px, const int c[5], bool trilinear) {
  auto vox = [&](int lvl) { return (size_t)B * (pz >> lvl) * (py >> lvl) * (px >> lvl); };
  const int f = trilinear ? 2 : 1;
  WsLayout L{};
  size_t off = 0;
  auto take = [&](size_t n) {
    size_t o = off;
    off += (n + 127) / 128 * 128;  // 256-byte alignment for bf16, more for fp32
    return o;
  };
  // normalised input as interleaved bf16 (hi, lo) pairs, rows padded to px + 8 voxels
  L.xhi = take((size_t)B * pz * py * (px + 8) * 2);
  L.xlo = 0;
  L.a0 = take(vox(0) * c[0]);        // inc.0 output; re-used for up4.0 output (A0 is dead by then)
  L.cat4 = take(vox(0) * 2 * c[0]);  // [x1 | up(u3)]
  L.p1 = take(vox(1) * c[0]);
  L.d1a = take(vox(1) * c[1]);
  L.cat3 = take(vox(1) * 2 * c[1]);  // [x2 | up(u2)]
  L.p2 = take(vox(2) * c[1]);
  L.d2a = take(vox(2) * c[2]);
  L.cat2 = take(vox(2) * 2 * c[2]);  // [x3 | up(u1)]
  L.p3 = take(vox(3) * c[2]);
  L.d3a = take(vox(3) * c[3]);
  L.cat1 = take(vox(3) * 2 * c[3]);  // [x4 | up(x5)]
  L.p4 = take(vox(4) * c[3]);
  L.d4a = take(vox(4) * (c[4] / f));
  L.x5 = take(vox(4) * (c[4] / f));
  L.u1a = take(vox(3) * (trilinear ? c[4] / 2 : c[3]));
  L.u1 = take(vox(3) * (c[3] / f));
  L.u2a = take(vox(2) * (trilinear ? c[3] / 2 : c[2]));
  L.u2 = take(vox(2) * (c[2] / f));
  L.u3a = take(vox(1) * (trilinear ? c[2] / 2 : c[1]));
  L.u3 = take(vox(1) * (c[1] / f));
  L.total = off;
  return L;
}
}  // namespace

Status Engine::ensure_workspace(int batch, int pz, int py, int px) {
  const size_t esz = precision_ == EXA_PRECISION_BF16 ? 2 : 4;
  const size_t need = layout(batch, pz, py, px, chan_, trilinear_).total * esz;
  if (need > ws_bytes_) {
    free_dev(ws_);
    ws_ = nullptr;
    ws_bytes_ = 0;
    EXA_CUDA(cudaMalloc(&ws_, need));
    ws_bytes_ = need;
  }
  return Status::OK();
}

Status Engine::conv(const ConvLayer& L, const Act& in, const Act& out, const HeadParams* head,
                    const ConvRegion* region, const Act* pool_out, cudaStream_t s) {
  cur_tag_ = (int)(&L - layers_);
  if (precision_ == EXA_PRECISION_BF16) {
    const bool zf = use_zfold_ && L.w_zfold && conv_zfold_supported(in, L.cout, use_pair_);
    layer_kind_[cur_tag_] = zf ? (use_pair_ ? 3 : 2) : 1;
    {
      Scope sc(this, CAT_CONV, s);
      if (zf) {
        // z-folded kernel: optional output sub-box and fused 2x2x2 max-pool
        const bool fuse_pool = pool_out && in.D % 2 == 0;
        EXA_TRY(launch_conv_zfold(in, out, L.w_zfold, L.bias, head, region,
                                  fuse_pool ? pool_out : nullptr, num_sms_, use_pair_, s));
        if (fuse_pool) return Status::OK();
      } else {
        EXA_TRY(launch_conv_umma(in, out, L.w_bf16, L.bias, head, num_sms_, s));
      }
    }
    if (pool_out) {
      Scope sc(this, CAT_POOL, s);
      EXA_TRY(launch_maxpool(out, *pool_out, s));
    }
    return Status::OK();
  }
  {
    Scope sc(this, CAT_CONV, s);
    EXA_TRY(launch_conv_fp32(in, out, L.w_f32, L.bias, s));
  }
  if (head) {
    Scope sc(this, CAT_HEAD, s);
    EXA_TRY(launch_head(out, *head, s));
  }
  if (pool_out) {
    Scope sc(this, CAT_POOL, s);
    EXA_TRY(launch_maxpool(out, *pool_out, s));
  }
  return Status::OK();
}

// One wave of `batch` patches through the whole network (unet3d.py:77-105).
Status Engine::run_network(const PatchSource& src, int batch, int pz, int py, int px,
                           const HeadParams& head, cudaStream_t s) {
  EXA_CHECK(finalized_, "weights not finalised");
  EXA_TRY(ensure_workspace(batch, pz, py, px));
  const bool f32 = precision_ == EXA_PRECISION_FP32;
  const size_t esz = f32 ? 4 : 2;
  const WsLayout L = layout(batch, pz, py, px, chan_, trilinear_);
  auto act = [&](size_t off, int lvl, int C, int cstride, int coff) {
    Act a;
    a.ptr = (char*)ws_ + off * esz;
    a.B = batch;
    a.D = pz >> lvl;
    a.H = py >> lvl;
    a.W = px >> lvl;
    a.C = C;
    a.cstride = cstride;
    a.coff = coff;
    a.fp32 = f32;
    return a;
  };
  // channel counts (unet3d.py:56-74); at every level the concat buffer is [skip c_k | upsampled c_k]
  const int c0 = chan_[0], c1 = chan_[1], c2 = chan_[2], c3 = chan_[3], c4 = chan_[4];
  const int fdiv = trilinear_ ? 2 : 1;
  const int cb = c4 / fdiv;                                   // bottleneck width
  const int m1 = trilinear_ ? c4 / 2 : c3, o1 = c3 / fdiv;    // up1 DoubleConv: mid, out
  const int m2 = trilinear_ ? c3 / 2 : c2, o2 = c2 / fdiv;
  const int m3 = trilinear_ ? c2 / 2 : c1, o3 = c1 / fdiv;
  const Act a0 = act(L.a0, 0, c0, c0, 0);
  const Act x1 = act(L.cat4, 0, c0, 2 * c0, 0), up4_slot = act(L.cat4, 0, c0, 2 * c0, c0),
            cat4 = act(L.cat4, 0, 2 * c0, 2 * c0, 0);
  const Act p1 = act(L.p1, 1, c0, c0, 0), d1a = act(L.d1a, 1, c1, c1, 0);
  const Act x2 = act(L.cat3, 1, c1, 2 * c1, 0), up3_slot = act(L.cat3, 1, c1, 2 * c1, c1),
            cat3 = act(L.cat3, 1, 2 * c1, 2 * c1, 0);
  const Act p2 = act(L.p2, 2, c1, c1, 0), d2a = act(L.d2a, 2, c2, c2, 0);
  const Act x3 = act(L.cat2, 2, c2, 2 * c2, 0), up2_slot = act(L.cat2, 2, c2, 2 * c2, c2),
            cat2 = act(L.cat2, 2, 2 * c2, 2 * c2, 0);
  const Act p3 = act(L.p3, 3, c2, c2, 0), d3a = act(L.d3a, 3, c3, c3, 0);
  const Act x4 = act(L.cat1, 3, c3, 2 * c3, 0), up1_slot = act(L.cat1, 3, c3, 2 * c3, c3),
            cat1 = act(L.cat1, 3, 2 * c3, 2 * c3, 0);
  const Act p4 = act(L.p4, 4, c3, c3, 0), d4a = act(L.d4a, 4, cb, cb, 0),
            x5 = act(L.x5, 4, cb, cb, 0);
  const Act u1a = act(L.u1a, 3, m1, m1, 0), u1 = act(L.u1, 3, o1, o1, 0);
  const Act u2a = act(L.u2a, 2, m2, m2, 0), u2 = act(L.u2, 2, o2, o2, 0);
  const Act u3a = act(L.u3a, 1, m3, m3, 0), u3 = act(L.u3, 1, o3, o3, 0);
  const Act u4a = a0;  // alias: A0 is dead after inc.3 (up4's mid width is c0 in both modes)
  // the last conv's activations before an unfused head (fp32 mode, or widths above 32)
  const Act u4 = act(L.cat4, 0, c0, c0, 0);  // cat4 is dead once up4.0 has run

  // x2 upsampling into the concat slot: trilinear interpolation (unet3d.py:248-250) or the
  // block's ConvTranspose3d (unet3d.py:254-256); `which` = 0..3 for up1..up4
  auto up = [&](int which, const Act& i, const Act& o, const ConvRegion* rg) {
    Scope sc(this, CAT_UPSAMPLE, s);
    if (trilinear_) return launch_upsample(i, o, rg, s);
    return launch_upconv(i, o, upconv_[which].w, upconv_[which].bias, rg, s);
  };
  // Only the inner [trim, P-trim) box of the last conv is ever read (inference.py:161-162), so
  // the last two convs and the last upsample are restricted to that box grown by their
  // receptive field.  Everything upstream feeds the pooled path and stays full.
  ConvRegion r_last, r_up40, r_ups;
  const ConvRegion *p_last = nullptr, *p_up40 = nullptr, *p_ups = nullptr;
  if (!f32 && head.trim > 0) {
    const int dims[3] = {pz, py, px};
    for (int i = 0; i < 3; ++i) {
      r_last.lo[i] = head.trim;
      r_last.hi[i] = dims[i] - head.trim;
      r_up40.lo[i] = std::max(r_last.lo[i] - 1, 0);
      r_up40.hi[i] = std::min(r_last.hi[i] + 1, dims[i]);
      r_ups.lo[i] = std::max(r_up40.lo[i] - 1, 0);
      r_ups.hi[i] = std::min(r_up40.hi[i] + 1, dims[i]);
    }
    p_last = &r_last;
    p_up40 = &r_up40;
    p_ups = &r_ups;
  }

  // the stem kernels write 32 channels per pass into channel slots of A0
  const int stem_groups = c0 / 32;
  if (!f32 && use_tc_stem_ && pz % 16 == 0 && py % 8 == 0 && px % 4 == 0) {
    // inc.0 on the tensor cores: gather/normalise into bf16 (hi, lo) pairs, then the Toeplitz-form conv
    __nv_bfloat16* xs = (__nv_bfloat16*)((char*)ws_ + L.xhi * esz);
    {
      Scope sc(this, CAT_STEM, s);
      EXA_TRY(launch_stem_split(src, batch, pz, py, px, xs, s));
    }
    for (int g = 0; g < stem_groups; ++g) {
      Scope sc(this, CAT_STEM, s);
      EXA_TRY(launch_stem_tc(xs, stem_band_ + (size_t)g * 9 * 128 * 16, stem_bias_ + g * 32,
                             act(L.a0, 0, 32, c0, g * 32), num_sms_, s));
    }
  } else {
    for (int g = 0; g < stem_groups; ++g) {
      Scope sc(this, CAT_STEM, s);
      EXA_TRY(launch_stem(src, stem_[g], act(L.a0, 0, 32, c0, g * 32), s));  // inc.0 (+gather/normalise)
    }
  }
  // the 1x1x1 head rides in the last conv's epilogue when that conv has 32 channels (bf16 mode)
  const bool fused_head = !f32 && c0 == 32;
  EXA_TRY(conv(layers_[1], a0, x1, nullptr, nullptr, &p1, s));   // inc.3 -> skip slot of CAT4 (+pool)
  EXA_TRY(conv(layers_[2], p1, d1a, nullptr, nullptr, nullptr, s));
  EXA_TRY(conv(layers_[3], d1a, x2, nullptr, nullptr, &p2, s));
  EXA_TRY(conv(layers_[4], p2, d2a, nullptr, nullptr, nullptr, s));
  EXA_TRY(conv(layers_[5], d2a, x3, nullptr, nullptr, &p3, s));
  EXA_TRY(conv(layers_[6], p3, d3a, nullptr, nullptr, nullptr, s));
  EXA_TRY(conv(layers_[7], d3a, x4, nullptr, nullptr, &p4, s));
  EXA_TRY(conv(layers_[8], p4, d4a, nullptr, nullptr, nullptr, s));
  EXA_TRY(conv(layers_[9], d4a, x5, nullptr, nullptr, nullptr, s));
  EXA_TRY(up(0, x5, up1_slot, nullptr));                    // cat([x4, up(x5)])  unet3d.py:288
  EXA_TRY(conv(layers_[10], cat1, u1a, nullptr, nullptr, nullptr, s));
  EXA_TRY(conv(layers_[11], u1a, u1, nullptr, nullptr, nullptr, s));
  EXA_TRY(up(1, u1, up2_slot, nullptr));
  EXA_TRY(conv(layers_[12], cat2, u2a, nullptr, nullptr, nullptr, s));
  EXA_TRY(conv(layers_[13], u2a, u2, nullptr, nullptr, nullptr, s));
  EXA_TRY(up(2, u2, up3_slot, nullptr));
  EXA_TRY(conv(layers_[14], cat3, u3a, nullptr, nullptr, nullptr, s));
  EXA_TRY(conv(layers_[15], u3a, u3, nullptr, nullptr, nullptr, s));
  EXA_TRY(up(3, u3, up4_slot, p_ups));
  EXA_TRY(conv(layers_[16], cat4, u4a, nullptr, p_up40, nullptr, s));
  if (fused_head || f32) {
    EXA_TRY(conv(layers_[17], u4a, u4, &head, p_last, nullptr, s));  // + 1x1x1 head, sigmoid, trim
  } else {
    EXA_TRY(conv(layers_[17], u4a, u4, nullptr, p_last, nullptr, s));
    Scope sc(this, CAT_HEAD, s);
    EXA_TRY(launch_head(u4, head, s));
  }
  return Status::OK();
}

// ---------------------------------------------------------------------------
// Engine: operator-level forward (replaces UNet3D.forward, unet3d.py:77-105)
// ---------------------------------------------------------------------------
Status Engine::forward(const float* x, float* logits, int batch, const int32_t patch[3],
                       cudaStream_t s) {
  EXA_CUDA(cudaSetDevice(device_));
  EXA_CHECK(x && logits && batch > 0, "forward: bad arguments");
  for (int i = 0; i < 3; ++i)
    EXA_CHECK(patch[i] > 0 && patch[i] % 16 == 0, "forward: patch dims must be multiples of 16");
  PatchSource src;
  src.x = x;
  HeadParams head;
  head.w = head_w_;
  head.b = head_b_;
  head.w_host = head_w_host_.data();
  head.b_host = head_b_host_.data();
  head.out = logits;
  head.C = out_channels_;
  head.trim = 0;
  head.apply_sigmoid = 0;
  return run_network(src, batch, patch[0], patch[1], patch[2], head, s);
}

// ---------------------------------------------------------------------------
// Engine: volume driver (replaces predict, inference.py:29-126)
// ---------------------------------------------------------------------------
Status Engine::histogram(const uint16_t* vol_dev, int64_t n, int clip, uint64_t* hist_dev,
                         cudaStream_t s) {
  EXA_CUDA(cudaSetDevice(device_));
  EXA_CHECK(vol_dev && hist_dev && n >= 0, "histogram: bad arguments");
  Scope sc(this, CAT_HIST, s);
  return launch_histogram(vol_dev, (size_t)n, clip, (unsigned long long*)hist_dev, s);
}

// LUT of (min(v,clip) - mn) / (mx - mn + 1e-8) clipped to [0,1], evaluated in float64 and
// rounded to float32 exactly where the reference does (img_util.py:527-531; the cast happens
// on assignment into the float32 batch array, inference.py:188-191).
Status Engine::set_normalization(double mn, double mx, int clip) {
  EXA_CUDA(cudaSetDevice(device_));
  EXA_CHECK(clip >= 0 && clip <= 65535, "brightness_clip must be in [0, 65535]");
  std::vector<float> lut(clip + 1);
  const double den = mx - mn + 1e-8;
  for (int v = 0; v <= clip; ++v) {
    double t = ((double)v - mn) / den;
    t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
    lut[v] = (float)t;
  }
  if (lut_clip_ != clip) {
    free_dev(lut_);
    lut_ = nullptr;
    EXA_CUDA(cudaMalloc(&lut_, sizeof(float) * (clip + 1)));
    lut_clip_ = clip;
  }
  EXA_CUDA(cudaMemcpy(lut_, lut.data(), sizeof(float) * (clip + 1), cudaMemcpyHostToDevice));
  norm_set_ = true;
  return Status::OK();
}

// The same table for a rank-compressed float volume (float_volume.cu): entry v is the normalised
// value of the v-th distinct clipped intensity, evaluated in float64 like img_util.py:527-531
Status Engine::set_normalization_table(const double* values, int n, double mn, double mx) {
  EXA_CUDA(cudaSetDevice(device_));
  EXA_CHECK(values && n >= 1 && n <= 65536, "set_normalization_table: 1..65536 values");
  const int clip = n - 1;
  std::vector<float> lut(n);
  const double den = mx - mn + 1e-8;
  for (int v = 0; v < n; ++v) {
    double t = (values[v] - mn) / den;
    t = t < 0.0 ? 0.0 : (t > 1.0 ? 1.0 : t);
    lut[v] = (float)t;
  }
  if (lut_clip_ != clip) {
    free_dev(lut_);
    lut_ = nullptr;
    EXA_CUDA(cudaMalloc(&lut_, sizeof(float) * n));
    lut_clip_ = clip;
  }
  EXA_CUDA(cudaMemcpy(lut_, lut.data(), sizeof(float) * n, cudaMemcpyHostToDevice));
  norm_set_ = true;
  return Status::OK();
}

Status Engine::slab_run(const uint16_t* slab_dev, int D, int H, int W, const exa_predict_params& p,
                        int row_begin, int row_end, cudaStream_t s) {
  EXA_CUDA(cudaSetDevice(device_));
  EXA_CHECK(finalized_, "weights not finalised");
  EXA_CHECK(norm_set_, "normalisation not set (exa_set_normalization)");
  EXA_CHECK(lut_clip_ == p.brightness_clip, "normalisation was set for a different brightness_clip");
  EXA_TRY(make_plan(D, H, W, p, &plan_));
  EXA_TRY(plan_slab(plan_, row_begin, row_end, &slab_));
  row_begin_ = row_begin;
  row_end_ = row_end;
  job_ready_ = false;
  const int rows = row_end - row_begin;
  const int per_row = plan_.ay.n * plan_.ax.n;
  const int n_slab = rows * per_row;
  if (n_slab == 0) {
    job_ready_ = true;
    return Status::OK();
  }
  EXA_CHECK(slab_dev != nullptr, "slab_run: null volume");
  const int t = p.trim;
  const size_t per_patch =
      (size_t)out_channels_ * (p.patch[0] - 2 * t) * (p.patch[1] - 2 * t) * (p.patch[2] - 2 * t);
  const size_t need = per_patch * n_slab * sizeof(float);
  if (need > probs_bytes_) {
    free_dev(probs_);
    probs_ = nullptr;
    probs_bytes_ = 0;
    EXA_CUDA(cudaMalloc(&probs_, need));
    probs_bytes_ = need;
  }
  // patch starts of this slab in the reference's order (inference.py:394-397)
  std::vector<int>& starts = starts_host_;  // member: must outlive the async copy below
  starts.resize((size_t)n_slab * 3);
  {
    size_t i = 0;
    for (int kz = row_begin; kz < row_end; ++kz)
      for (int ky = 0; ky < plan_.ay.n; ++ky)
        for (int kx = 0; kx < plan_.ax.n; ++kx) {
          starts[i++] = kz * plan_.az.stride;
          starts[i++] = ky * plan_.ay.stride;
          starts[i++] = kx * plan_.ax.stride;
        }
  }
  if (starts.size() > starts_cap_) {
    free_dev(starts_dev_);
    starts_dev_ = nullptr;
    starts_cap_ = 0;
    EXA_CUDA(cudaMalloc(&starts_dev_, starts.size() * sizeof(int)));
    starts_cap_ = starts.size();
  }
  EXA_CUDA(cudaMemcpyAsync(starts_dev_, starts.data(), starts.size() * sizeof(int),
                           cudaMemcpyHostToDevice, s));
  // (pageable source: the runtime stages the data before cudaMemcpyAsync returns)

  int batch = p.batch > 0 ? p.batch : 32;
  batch = std::min(batch, n_slab);
  for (int i0 = 0; i0 < n_slab; i0 += batch) {
    const int nb = std::min(batch, n_slab - i0);
    PatchSource src;
    src.vol = slab_dev;
    src.gD = D;
    src.gH = H;
    src.gW = W;
    src.vz0 = slab_.in_z0;
    src.vD = slab_.in_z1 - slab_.in_z0;
    src.lut = lut_;
    src.clip = p.brightness_clip;
    src.starts = starts_dev_ + (size_t)i0 * 3;
    HeadParams head;
    head.w = head_w_;
    head.b = head_b_;
    head.w_host = head_w_host_.data();
    head.b_host = head_b_host_.data();
    head.out = probs_ + (size_t)i0 * per_patch;
    head.C = out_channels_;
    head.trim = t;
    head.apply_sigmoid = 1;  // inference.py:158
    EXA_TRY(run_network(src, nb, p.patch[0], p.patch[1], p.patch[2], head, s));
    if (progress_cb_) {
      progress_done_ += nb;
      struct Rec {
        exa_progress_fn cb;
        void* user;
        int64_t done, total;
      };
      Rec* rec = new Rec{progress_cb_, progress_user_, progress_done_,
                         std::max(progress_total_, progress_done_)};
      EXA_CUDA(cudaLaunchHostFunc(s, [](void* q) {
        Rec* r = static_cast<Rec*>(q);
        r->cb(r->user, r->done, r->total);
        delete r;
      }, rec));
    }
    if (band_.on) {
      job_ready_ = true;  // the stitch below only reads patches that are complete
      EXA_TRY(stream_band(i0 + nb, n_slab, s));
      job_ready_ = false;
    }
  }
  job_ready_ = true;
  return Status::OK();
}

// Patches [0, patches_done) of the current one-row slab job are finished: every output row
// y < ky_done*stride + trim is covered only by finished patches (the next window row starts
// there), so those rows of the group's planes are final -- stitch them and start their D2H.
Status Engine::stream_band(int patches_done, int n_slab, cudaStream_t s) {
  const AxisGeom& ay = plan_.ay;
  const int ky_done = patches_done / plan_.ax.n;
  const int y_done = patches_done >= n_slab ? ay.dim : std::min(ky_done * ay.stride + ay.trim, ay.dim);
  if (y_done <= band_.y_done || band_.z1 <= band_.z0) return Status::OK();
  const size_t plane = (size_t)plan_.H * plan_.W;
  EXA_TRY(stitch_planes(band_.seed, band_.out_dev + (size_t)(band_.z0 - band_.out_zbase) * plane,
                        band_.out_cstride, band_.z0, band_.z1, s, band_.y_done, y_done));
  if (band_.out_host)
    EXA_TRY(copy_planes_to_host(band_.out_dev, band_.out_cstride, band_.out_host, band_.host_cstride,
                                band_.out_zbase, band_.z0, band_.z1, plane, s, band_.y_done, y_done));
  band_.y_done = y_done;
  return Status::OK();
}

static StitchArgs stitch_base(const Plan& plan, const float* probs, int C, int row_begin,
                              int row_end) {
  StitchArgs a;
  a.probs = probs;
  a.C = C;
  a.az = plan.az;
  a.ay = plan.ay;
  a.ax = plan.ax;
  a.row_begin = row_begin;
  a.row_end = row_end;
  return a;
}

Status Engine::slab_partial(float* halo_dev, cudaStream_t s) {
  EXA_CUDA(cudaSetDevice(device_));
  EXA_CHECK(job_ready_, "slab_partial: no slab job (call exa_slab_run first)");
  const int nz = slab_.halo_z1 - slab_.halo_z0;
  if (nz <= 0) return Status::OK();
  EXA_CHECK(halo_dev != nullptr, "slab_partial: null buffer");
  StitchArgs a = stitch_base(plan_, probs_, out_channels_, row_begin_, row_end_);
  a.z_begin = slab_.halo_z0;
  a.z_end = slab_.halo_z1;
  a.out = halo_dev;
  a.out_cstride = (size_t)nz * plan_.H * plan_.W;
  a.finalize = 0;
  Scope sc(this, CAT_STITCH, s);
  return launch_stitch(a, s);
}

Status Engine::slab_stitch(const float* seed_dev, float* out_dev, int64_t channel_stride,
                           cudaStream_t s) {
  EXA_CUDA(cudaSetDevice(device_));
  EXA_CHECK(job_ready_, "slab_stitch: no slab job (call exa_slab_run first)");
  const int nz = slab_.out_z1 - slab_.out_z0;
  if (nz <= 0) return Status::OK();
  const size_t dense = (size_t)nz * plan_.H * plan_.W;
  EXA_CHECK(channel_stride == 0 || (size_t)channel_stride >= dense, "slab_stitch: channel stride too small");
  return stitch_planes(seed_dev, out_dev, channel_stride ? (size_t)channel_stride : dense,
                       slab_.out_z0, slab_.out_z1, s);
}

// finished planes [z0, z1) of the current slab job -> out (plane z0 first), channel stride given
Status Engine::stitch_planes(const float* seed_dev, float* out_dev, size_t out_cstride, int z0,
                             int z1, cudaStream_t s, int y0, int y1) {
  EXA_CHECK(job_ready_, "stitch: no slab job");
  if (z1 <= z0) return Status::OK();
  EXA_CHECK(out_dev != nullptr, "slab_stitch: null buffer");
  StitchArgs a = stitch_base(plan_, probs_, out_channels_, row_begin_, row_end_);
  a.z_begin = z0;
  a.z_end = z1;
  a.y_begin = y0;
  a.y_end = y1;
  a.out = out_dev;
  a.out_cstride = out_cstride;
  a.finalize = 1;
  if (seed_dev != nullptr && slab_.seed_z1 > slab_.seed_z0) {
    a.seed = seed_dev;
    a.seed_z0 = slab_.seed_z0;
    a.seed_z1 = slab_.seed_z1;
  }
  const bool to_peers = peer_local_ && !peer_bases_.empty() && out_dev >= peer_local_ &&
                        out_dev < peer_local_ + peer_elems_;
  if (to_peers && !peer_ce_) {
    const ptrdiff_t off = out_dev - peer_local_;
    a.n_peers = (int)peer_bases_.size();
    for (int i = 0; i < a.n_peers; ++i) a.peer_out[i] = peer_bases_[i] + off;
  }
  {
    Scope sc(this, CAT_STITCH, s);
    EXA_TRY(launch_stitch(a, s));
  }
  if (to_peers && peer_ce_) EXA_TRY(copy_planes_to_peers(out_dev, out_cstride, z1 - z0, y0, y1, s));
  return Status::OK();
}

// Copy-engine form of the fused gather: rows [y0, y1) (all rows when y1 <= y0) of the nz planes at
// out_dev, every channel, to the same offsets of every peer's copy.  The transfers run on side
// streams, ordered after the stitch just queued on s; join_peer_copies() orders s after them.
Status Engine::copy_planes_to_peers(const float* out_dev, size_t out_cstride, int nz, int y0, int y1,
                                    cudaStream_t s) {
  if (!peer_ready_) {
    const char* ns = getenv("EXA_PEER_STREAMS");
    if (ns && atoi(ns) >= 1 && atoi(ns) <= kPeerStreams) n_peer_streams_ = atoi(ns);
    EXA_CUDA(cudaEventCreateWithFlags(&peer_ready_, cudaEventDisableTiming));
    for (int i = 0; i < kPeerStreams; ++i) {
      EXA_CUDA(cudaStreamCreateWithFlags(&peer_stream_[i], cudaStreamNonBlocking));
      EXA_CUDA(cudaEventCreateWithFlags(&peer_done_[i], cudaEventDisableTiming));
    }
  }
  const size_t plane = (size_t)plan_.H * plan_.W;
  const ptrdiff_t off = out_dev - peer_local_;
  const bool band = y1 > y0 && !(y0 == 0 && (size_t)y1 * plan_.W == plane);
  EXA_CUDA(cudaEventRecord(peer_ready_, s));
  for (int i = 0; i < n_peer_streams_; ++i) EXA_CUDA(cudaStreamWaitEvent(peer_stream_[i], peer_ready_, 0));
  for (size_t i = 0; i < peer_bases_.size(); ++i) {
    cudaStream_t ps = peer_stream_[i % n_peer_streams_];
    for (int c = 0; c < out_channels_; ++c) {
      const float* src = out_dev + c * out_cstride;
      float* dst = peer_bases_[i] + off + c * out_cstride;
      if (!band) {
        EXA_CUDA(cudaMemcpyAsync(dst, src, (size_t)nz * plane * 4, cudaMemcpyDeviceToDevice, ps));
      } else {
        const size_t yoff = (size_t)y0 * plan_.W;
        EXA_CUDA(cudaMemcpy2DAsync(dst + yoff, plane * 4, src + yoff, plane * 4,
                                   (size_t)(y1 - y0) * plan_.W * 4, (size_t)nz,
                                   cudaMemcpyDeviceToDevice, ps));
      }
    }
  }
  peer_pending_ = true;
  return Status::OK();
}

Status Engine::join_peer_copies(cudaStream_t s) {
  if (!peer_pending_) return Status::OK();
  for (int i = 0; i < n_peer_streams_; ++i) {
    EXA_CUDA(cudaEventRecord(peer_done_[i], peer_stream_[i]));
    EXA_CUDA(cudaStreamWaitEvent(s, peer_done_[i], 0));
  }
  peer_pending_ = false;
  return Status::OK();
}

Status Engine::set_peer_outputs(float* local_base, int64_t elems, float* const* peer_bases,
                                int n_peers) {
  EXA_CHECK(n_peers >= 0 && n_peers <= EXA_MAX_PEERS, "set_peer_outputs: too many peers");
  EXA_CHECK(n_peers == 0 || (local_base && peer_bases && elems > 0), "set_peer_outputs: null argument");
  peer_bases_.clear();
  peer_local_ = n_peers > 0 ? local_base : nullptr;
  peer_elems_ = n_peers > 0 ? elems : 0;
  {
    // default: copy-engine transfers; EXA_GATHER=store: stores from the stitch kernel (measured
    // 7 ms slower per 1024^3 step on 8 GPUs: profiles/r02i_*)
    const char* g = getenv("EXA_GATHER");
    peer_ce_ = !(g && std::string(g) == "store");
  }
  for (int i = 0; i < n_peers; ++i) {
    EXA_CHECK(peer_bases[i] != nullptr, "set_peer_outputs: null peer pointer");
    peer_bases_.push_back(peer_bases[i]);
  }
  return Status::OK();
}

// Rows [R0, R1) of the volume in groups of z patch-rows.  Each group is a slab job: run its
// patches, stitch the planes it owns (seeded with the previous group's partial sums for the shared
// planes, so the fp32 summation order is the reference's), hand partial sums to the next group.
// Finished planes are copied to the host on a second stream while the next group computes.
// vol_dev points at plane vol_z0; out_dev / out_host at plane out_zbase (channel strides in
// elements).  The last group's partial sums for the next rank go to halo_out.  With defer_seed the
// planes of the first group that need the previous rank's partial sums are left out (their
// patches stay in a held buffer) until pipeline_finish() gets that seed.
Status Engine::pipeline_rows(const uint16_t* vol_dev, int vol_z0, int D, int H, int W,
                             const exa_predict_params& p, int R0, int R1, float* out_dev,
                             size_t out_cstride, float* out_host, size_t host_cstride,
                             int out_zbase, float* halo_out, bool defer_seed, cudaStream_t s) {
  Plan plan;
  EXA_TRY(make_plan(D, H, W, p, &plan));
  EXA_CHECK(R0 >= 0 && R0 <= R1 && R1 <= plan.az.n, "row range out of bounds");
  const size_t plane = (size_t)H * W;
  held_valid_ = false;
  // group size: at least one wave of patches per group; a single group when planes can be
  // covered by more than two rows (no pairwise hand-over possible)
  const int per_row = plan.ay.n * plan.ax.n;
  const int batch = p.batch > 0 ? p.batch : 32;
  int rows_per_group = std::max(1, ceil_div(batch, per_row));
  const int keep = plan.az.patch - 2 * plan.az.trim;
  if (keep > 2 * plan.az.stride) rows_per_group = plan.az.n;
  if (out_host && !copy_stream_) {
    EXA_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
  }
  const size_t seed_elems = (size_t)out_channels_ * keep * plane;  // upper bound on shared planes
  if (rows_per_group < R1 - R0 && seed_elems * sizeof(float) > seed_bytes_) {
    free_dev(seed_);
    seed_ = nullptr;
    seed_bytes_ = 0;
    EXA_CUDA(cudaMalloc(&seed_, seed_elems * sizeof(float)));
    seed_bytes_ = seed_elems * sizeof(float);
  }
  // EXA_STITCH_OVERLAP=1: the stitch of a group (and its peer stores, when the gather is fused
  // into it) runs on a second stream while the next group's convolutions run on s; two patch
  // buffers alternate, a buffer is rewritten only after the stitch that read it.  Off by
  // default: measured on the same boxes it is neutral at 1 GPU and slower at 2 and 8 GPUs
  // (the stores compete with the HBM-bound stem of the next group).
  if (!stitch_stream_) {
    int lo = 0, hi = 0;
    EXA_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    EXA_CUDA(cudaStreamCreateWithPriority(&stitch_stream_, cudaStreamNonBlocking, hi));
    EXA_CUDA(cudaEventCreateWithFlags(&conv_done_, cudaEventDisableTiming));
    EXA_CUDA(cudaEventCreateWithFlags(&probs_free_, cudaEventDisableTiming));
    EXA_CUDA(cudaEventCreateWithFlags(&alt_free_, cudaEventDisableTiming));
  }
  cudaStream_t ss = overlap_stitch_ ? stitch_stream_ : s;
  Status st = Status::OK();
  bool have_seed = false;
  bool probs_busy = false, alt_busy = false;  // a stitch on ss still reads that buffer
  cudaEvent_t last_rec = nullptr;             // last event recorded on ss
  for (int r0 = R0; r0 < R1 && st.ok; r0 += rows_per_group) {
    const int r1 = std::min(r0 + rows_per_group, R1);
    if (probs_busy) EXA_CUDA(cudaStreamWaitEvent(s, probs_free_, 0));
    probs_busy = false;
    // one-row groups with a host destination stream finished y-bands while the row still runs
    band_ = BandSink();
    const bool ce_peers = peer_ce_ && peer_local_ && !peer_bases_.empty() && out_dev >= peer_local_ &&
                          out_dev < peer_local_ + peer_elems_;
    if ((out_host || ce_peers) && r1 - r0 == 1 && ss == s) {
      exa_slab_plan g;
      st = plan_slab(plan, r0, r1, &g);
      if (!st.ok) break;
      band_.on = true;
      band_.seed = have_seed ? seed_ : nullptr;
      band_.out_dev = out_dev;
      band_.out_cstride = out_cstride;
      band_.out_host = out_host;
      band_.host_cstride = host_cstride;
      band_.out_zbase = out_zbase;
      band_.z0 = (r0 == R0 && defer_seed && g.seed_z1 > g.seed_z0) ? g.seed_z1 : g.out_z0;
      band_.z1 = g.out_z1;
    }
    st = slab_run(vol_dev + (size_t)(r0 * plan.az.stride - vol_z0) * plane, D, H, W, p, r0, r1, s);
    const bool banded = band_.on;
    band_.on = false;
    if (!st.ok) break;
    if (ss != s) {
      EXA_CUDA(cudaEventRecord(conv_done_, s));
      EXA_CUDA(cudaStreamWaitEvent(ss, conv_done_, 0));
    }
    const bool first = r0 == R0, last = r1 == R1;
    const bool hold = first && defer_seed && slab_.seed_z1 > slab_.seed_z0;
    const int z0 = hold ? slab_.seed_z1 : slab_.out_z0, z1 = slab_.out_z1;
    if (z1 > z0 && !banded) {
      st = stitch_planes(have_seed ? seed_ : nullptr, out_dev + (size_t)(z0 - out_zbase) * plane,
                         out_cstride, z0, z1, ss);
      if (!st.ok) break;
    }
    have_seed = slab_.halo_z1 > slab_.halo_z0;
    if (have_seed) {  // after the stitch above has consumed the previous seed
      if (last && halo_out == nullptr) {
        st = Status::Err("slab_predict: rows end before the volume does but no halo buffer given");
        break;
      }
      st = slab_partial(last ? halo_out : seed_, ss);
      if (!st.ok) break;
      if (last) have_seed = false;
    }
    if (out_host && z1 > z0 && !banded) {
      st = copy_planes_to_host(out_dev, out_cstride, out_host, host_cstride, out_zbase, z0, z1,
                               plane, ss);
      if (!st.ok) break;
    }
    if (ss != s) {
      EXA_CUDA(cudaEventRecord(probs_free_, ss));
      probs_busy = true;
      last_rec = probs_free_;
    }
    if (hold) {  // these patches wait for the previous rank's seed; they leave the rotation
      held_slab_ = slab_;
      held_plan_ = plan_;
      held_rows_[0] = r0;
      held_rows_[1] = r1;
      std::swap(probs_, probs_hold_);
      std::swap(probs_bytes_, probs_hold_bytes_);
      held_valid_ = true;
      job_ready_ = false;
      probs_busy = false;  // the buffer now in probs_ has no reader (pipeline_finish orders s after ss)
    } else if (ss != s) {
      std::swap(probs_, probs_alt_);
      std::swap(probs_bytes_, probs_alt_bytes_);
      std::swap(probs_free_, alt_free_);
      std::swap(probs_busy, alt_busy);
      job_ready_ = false;  // slab_ no longer describes probs_
    }
  }
  // everything queued on the stitch stream is ordered before later work on s
  if (last_rec) EXA_CUDA(cudaStreamWaitEvent(s, last_rec, 0));
  return st;
}

// D2H of finished planes [z0, z1) on the copy stream, ordered after the work queued on s so far
Status Engine::copy_planes_to_host(const float* out_dev, size_t out_cstride, float* out_host,
                                   size_t host_cstride, int out_zbase, int z0, int z1, size_t plane,
                                   cudaStream_t s, int y0, int y1) {
  if (!copy_event_) EXA_CUDA(cudaEventCreateWithFlags(&copy_event_, cudaEventDisableTiming));
  EXA_CUDA(cudaEventRecord(copy_event_, s));
  EXA_CUDA(cudaStreamWaitEvent(copy_stream_, copy_event_, 0));
  const bool band = y1 > y0 && !(y0 == 0 && (size_t)y1 * plan_.W == plane);
  for (int c = 0; c < out_channels_; ++c) {
    const size_t zoff = (size_t)(z0 - out_zbase) * plane;
    if (!band) {
      EXA_CUDA(cudaMemcpyAsync(out_host + c * host_cstride + zoff, out_dev + c * out_cstride + zoff,
                               (size_t)(z1 - z0) * plane * 4, cudaMemcpyDeviceToHost, copy_stream_));
    } else {  // rows [y0, y1) of every plane: one contiguous run per plane, plane pitch apart
      const size_t yoff = (size_t)y0 * plan_.W;
      EXA_CUDA(cudaMemcpy2DAsync(out_host + c * host_cstride + zoff + yoff, plane * 4,
                                 out_dev + c * out_cstride + zoff + yoff, plane * 4,
                                 (size_t)(y1 - y0) * plan_.W * 4, (size_t)(z1 - z0),
                                 cudaMemcpyDeviceToHost, copy_stream_));
    }
  }
  return Status::OK();
}

// Second half of a deferred pipeline: stitch the held planes with the previous rank's partial
// sums, copy them out and wait for every copy of the job.
Status Engine::pipeline_finish(const float* seed_in, float* out_dev, size_t out_cstride,
                               float* out_host, size_t host_cstride, int out_zbase,
                               cudaStream_t s) {
  Status st = Status::OK();
  if (held_valid_) {
    std::swap(probs_, probs_hold_);
    std::swap(probs_bytes_, probs_hold_bytes_);
    slab_ = held_slab_;
    plan_ = held_plan_;
    row_begin_ = held_rows_[0];
    row_end_ = held_rows_[1];
    job_ready_ = true;
    held_valid_ = false;
    const int z0 = slab_.seed_z0, z1 = slab_.seed_z1;
    const size_t plane = (size_t)plan_.H * plan_.W;
    if (seed_in == nullptr) {
      st = Status::Err("slab_finish: the first rows share planes with the previous rank: seed needed");
    } else {
      st = stitch_planes(seed_in, out_dev + (size_t)(z0 - out_zbase) * plane, out_cstride, z0, z1, s);
    }
    if (st.ok && out_host)
      st = copy_planes_to_host(out_dev, out_cstride, out_host, host_cstride, out_zbase, z0, z1,
                               plane, s);
  }
  if (out_host && copy_stream_) {
    cudaError_t e = cudaStreamSynchronize(copy_stream_);
    if (e != cudaSuccess && st.ok)
      st = Status::Err(std::string("predict: D2H failed: ") + cudaGetErrorString(e));
  }
  if (st.ok) st = join_peer_copies(s);   // the peers' copies are complete before later work on s
  return st;
}

Status Engine::predict_pipeline(const uint16_t* vol_dev, int D, int H, int W,
                                const exa_predict_params& p, float* out_dev, float* out_host,
                                cudaStream_t s) {
  Plan plan;
  EXA_TRY(make_plan(D, H, W, p, &plan));
  progress_done_ = 0;
  progress_total_ = plan.n_patches;
  const int bins = p.brightness_clip + 1;
  const size_t plane = (size_t)H * W, nvox = plane * D;
  if (!hist_dev_) EXA_CUDA(cudaMalloc(&hist_dev_, sizeof(unsigned long long) * 65536));
  EXA_TRY(histogram(vol_dev, (int64_t)nvox, p.brightness_clip, (uint64_t*)hist_dev_, s));
  std::vector<uint64_t> hist(bins);
  EXA_CUDA(cudaMemcpyAsync(hist.data(), hist_dev_, sizeof(uint64_t) * bins, cudaMemcpyDeviceToHost,
                           s));
  EXA_CUDA(cudaStreamSynchronize(s));
  double mn = 0, mx = 0;
  EXA_TRY(percentiles_from_hist(hist.data(), bins, p.pct_lo, p.pct_hi, &mn, &mx));
  EXA_TRY(set_normalization(mn, mx, p.brightness_clip));
  if (plan.n_patches == 0) {  // volume smaller than one stride: the reference returns zeros
    EXA_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(float) * out_channels_ * nvox, s));
    if (out_host) {
      EXA_CUDA(cudaMemcpyAsync(out_host, out_dev, sizeof(float) * out_channels_ * nvox,
                               cudaMemcpyDeviceToHost, s));
    }
    return Status::OK();
  }
  Status st = pipeline_rows(vol_dev, 0, D, H, W, p, 0, plan.az.n, out_dev, nvox, out_host, nvox, 0,
                            nullptr, false, s);
  Status fin = pipeline_finish(nullptr, out_dev, nvox, out_host, nvox, 0, s);
  return st.ok ? fin : st;
}

// One rank's rows of a sharded volume (normalisation already set): everything but the planes
// shared with the previous rank is finished and on its way to the host when this returns.
Status Engine::slab_predict(const uint16_t* slab_dev, int D, int H, int W,
                            const exa_predict_params& p, int row_begin, int row_end, float* out_dev,
                            int64_t channel_stride, float* out_host, int64_t host_channel_stride,
                            float* halo_dev, cudaStream_t s) {
  EXA_CUDA(cudaSetDevice(device_));
  Plan plan;
  EXA_TRY(make_plan(D, H, W, p, &plan));
  exa_slab_plan whole;
  EXA_TRY(plan_slab(plan, row_begin, row_end, &whole));
  pipe_zbase_ = whole.out_z0;
  if (row_begin == row_end) return Status::OK();
  const size_t dense = (size_t)std::max(whole.out_z1 - whole.out_z0, 0) * H * W;
  EXA_CHECK(slab_dev && (out_dev || dense == 0), "slab_predict: null buffer");
  EXA_CHECK((size_t)channel_stride >= dense && (!out_host || (size_t)host_channel_stride >= dense),
            "slab_predict: channel stride too small");
  return pipeline_rows(slab_dev, whole.in_z0, D, H, W, p, row_begin, row_end, out_dev,
                       (size_t)channel_stride, out_host, (size_t)host_channel_stride, whole.out_z0,
                       halo_dev, true, s);
}

Status Engine::slab_finish(const float* seed_dev, float* out_dev, int64_t channel_stride,
                           float* out_host, int64_t host_channel_stride, cudaStream_t s) {
  EXA_CUDA(cudaSetDevice(device_));
  Status st = pipeline_finish(seed_dev, out_dev, (size_t)channel_stride, out_host,
                              (size_t)host_channel_stride, pipe_zbase_, s);
  cudaError_t e = cudaStreamSynchronize(s);
  if (e != cudaSuccess && st.ok) st = Status::Err(std::string("slab_finish: ") + cudaGetErrorString(e));
  return st;
}

Status Engine::predict_device(const uint16_t* vol_dev, int D, int H, int W,
                              const exa_predict_params& p, float* out_dev, cudaStream_t s) {
  EXA_CUDA(cudaSetDevice(device_));
  EXA_CHECK(vol_dev && out_dev, "predict: null buffer");
  return predict_pipeline(vol_dev, D, H, W, p, out_dev, nullptr, s);
}

Status Engine::predict_host(const uint16_t* vol, int D, int H, int W, const exa_predict_params& p,
                            float* out) {
  EXA_CUDA(cudaSetDevice(device_));
  EXA_CHECK(vol && out, "predict: null buffer");
  EXA_CHECK(finalized_, "weights not finalised");
  EXA_TRY(check_params(p));
  EXA_CHECK(D > 0 && H > 0 && W > 0, "volume dims must be positive");
  const size_t nvox = (size_t)D * H * W;
  // device staging buffers are kept between calls (re-allocated only when they must grow)
  if (nvox * 2 + 16 > vol_stage_bytes_) {
    free_dev(vol_stage_);
    vol_stage_ = nullptr;
    vol_stage_bytes_ = 0;
    EXA_CUDA(cudaMalloc(&vol_stage_, nvox * 2 + 16));
    vol_stage_bytes_ = nvox * 2 + 16;
  }
  if (nvox * 4 * out_channels_ > out_stage_bytes_) {
    free_dev(out_stage_);
    out_stage_ = nullptr;
    out_stage_bytes_ = 0;
    EXA_CUDA(cudaMalloc(&out_stage_, nvox * 4 * out_channels_));
    out_stage_bytes_ = nvox * 4 * out_channels_;
  }
  cudaStream_t s = nullptr;
  EXA_CUDA(cudaMemcpyAsync(vol_stage_, vol, nvox * 2, cudaMemcpyHostToDevice, s));
  EXA_TRY(predict_pipeline((const uint16_t*)vol_stage_, D, H, W, p, (float*)out_stage_, out, s));
  EXA_CUDA(cudaStreamSynchronize(s));
  return Status::OK();
}

}  // namespace exa
