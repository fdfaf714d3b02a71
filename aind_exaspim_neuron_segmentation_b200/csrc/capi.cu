// C ABI (include/exaspim_b200.h) over exa::Engine.  Nothing throws across this boundary:
// every entry point converts Status / C++ exceptions into a negative code and a message
// retrievable with exa_last_error().
#include <new>
#include <string>

#include "../../include/exaspim_b200.h"
#include "engine.h"
#include "float_volume.h"
#include "trainer.h"
#include "watershed.h"

struct exa_trainer {
  exa::Trainer impl;
  exa_trainer(int device, int precision) : impl(device, precision) {}
};

struct exa_engine {
  exa::Engine impl;
  exa_engine(int device, int precision) : impl(device, precision) {}
};

namespace {

thread_local std::string g_create_error;

int code_for(const std::string& msg) {
  if (msg.find("cuda") != std::string::npos || msg.find("CUDA") != std::string::npos)
    return EXA_ERR_CUDA;
  if (msg.find("not finalised") != std::string::npos || msg.find("not set") != std::string::npos ||
      msg.find("no slab job") != std::string::npos)
    return EXA_ERR_STATE;
  return EXA_ERR_INVALID;
}

template <typename F>
int guarded(exa_engine* e, F&& f) {
  if (!e) return EXA_ERR_INVALID;
  try {
    exa::Status s = f();
    if (s.ok) return EXA_OK;
    e->impl.last_error = s.msg;
    return code_for(s.msg);
  } catch (const std::exception& ex) {
    e->impl.last_error = std::string("exception: ") + ex.what();
    return EXA_ERR_INVALID;
  } catch (...) {
    e->impl.last_error = "unknown exception";
    return EXA_ERR_INVALID;
  }
}

template <typename F>
int guarded_static(F&& f) {
  try {
    exa::Status s = f();
    if (s.ok) return EXA_OK;
    g_create_error = s.msg;
    return code_for(s.msg);
  } catch (const std::exception& ex) {
    g_create_error = std::string("exception: ") + ex.what();
    return EXA_ERR_INVALID;
  } catch (...) {
    g_create_error = "unknown exception";
    return EXA_ERR_INVALID;
  }
}

template <typename F>
int guarded_train(exa_trainer* t, F&& f) {
  if (!t) return EXA_ERR_INVALID;
  try {
    exa::Status s = f();
    if (s.ok) return EXA_OK;
    t->impl.last_error = s.msg;
    return code_for(s.msg);
  } catch (const std::exception& ex) {
    t->impl.last_error = std::string("exception: ") + ex.what();
    return EXA_ERR_INVALID;
  } catch (...) {
    t->impl.last_error = "unknown exception";
    return EXA_ERR_INVALID;
  }
}

}  // namespace

extern "C" {

const char* exa_version(void) { return "exaspim_b200 0.1 (sm_100a)"; }

int exa_create(int device, int precision, exa_engine** out) {
  if (!out) return EXA_ERR_INVALID;
  *out = nullptr;
  exa_engine* e = new (std::nothrow) exa_engine(device, precision);
  if (!e) {
    g_create_error = "out of host memory";
    return EXA_ERR_INVALID;
  }
  exa::Status s;
  try {
    s = e->impl.init();
  } catch (...) {
    s = exa::Status::Err("exception during init");
  }
  if (!s.ok) {
    g_create_error = s.msg;
    delete e;
    return code_for(s.msg);
  }
  *out = e;
  return EXA_OK;
}

int exa_destroy(exa_engine* e) {
  if (!e) return EXA_ERR_INVALID;
  delete e;
  return EXA_OK;
}

const char* exa_last_error(const exa_engine* e) {
  return e ? e->impl.last_error.c_str() : g_create_error.c_str();
}

int exa_load_weight(exa_engine* e, const char* name, const void* data, const int64_t* shape,
                    int ndim, int dtype) {
  return guarded(e, [&] {
    EXA_CHECK(name != nullptr, "load_weight: null name");
    EXA_CHECK(ndim == 0 || shape != nullptr, "load_weight: null shape");
    return e->impl.load_weight(name, data, shape, ndim, dtype);
  });
}

int exa_finalize_weights(exa_engine* e) {
  return guarded(e, [&] { return e->impl.finalize_weights(); });
}

int exa_out_channels(const exa_engine* e) { return e ? e->impl.out_channels() : EXA_ERR_INVALID; }

int exa_forward(exa_engine* e, const float* x, float* logits, int batch, const int32_t patch[3],
                void* stream) {
  return guarded(e, [&] {
    EXA_CHECK(patch != nullptr, "forward: null patch");
    return e->impl.forward(x, logits, batch, patch, (cudaStream_t)stream);
  });
}

int exa_predict(exa_engine* e, const uint16_t* vol, int D, int H, int W,
                const exa_predict_params* p, float* out) {
  return guarded(e, [&] {
    EXA_CHECK(p != nullptr, "predict: null params");
    return e->impl.predict_host(vol, D, H, W, *p, out);
  });
}

int exa_set_progress_callback(exa_engine* e, exa_progress_fn cb, void* user) {
  return guarded(e, [&] {
    e->impl.set_progress(cb, user);
    return exa::Status::OK();
  });
}

int exa_predict_device(exa_engine* e, const uint16_t* vol_dev, int D, int H, int W,
                       const exa_predict_params* p, float* out_dev, void* stream) {
  return guarded(e, [&] {
    EXA_CHECK(p != nullptr, "predict: null params");
    return e->impl.predict_device(vol_dev, D, H, W, *p, out_dev, (cudaStream_t)stream);
  });
}

int exa_plan_slab(int D, int H, int W, const exa_predict_params* p, int row_begin, int row_end,
                  exa_slab_plan* plan) {
  return guarded_static([&] {
    EXA_CHECK(p && plan, "plan_slab: null argument");
    exa::Plan pl;
    EXA_TRY(exa::make_plan(D, H, W, *p, &pl));
    return exa::plan_slab(pl, row_begin, row_end, plan);
  });
}

int exa_histogram(exa_engine* e, const uint16_t* vol_dev, int64_t n, int clip, uint64_t* hist_dev,
                  void* stream) {
  return guarded(e, [&] {
    return e->impl.histogram(vol_dev, n, clip, hist_dev, (cudaStream_t)stream);
  });
}

int exa_percentiles_from_hist(const uint64_t* hist, int nbins, double q_lo, double q_hi,
                              double* mn, double* mx) {
  return guarded_static([&] { return exa::percentiles_from_hist(hist, nbins, q_lo, q_hi, mn, mx); });
}

int exa_set_normalization(exa_engine* e, double mn, double mx, int clip) {
  return guarded(e, [&] { return e->impl.set_normalization(mn, mx, clip); });
}

int exa_set_normalization_table(exa_engine* e, const double* values, int n, double mn, double mx) {
  return guarded(e, [&] { return e->impl.set_normalization_table(values, n, mn, mx); });
}

int exa_compress_float_volume(const void* vol_dev, int is_double, int64_t n, double clip,
                              uint16_t* idx_dev, double* table_out, int* n_table, void* stream) {
  return guarded_static([&] {
    return exa::compress_float_volume(vol_dev, is_double, n, clip, idx_dev, table_out, n_table,
                                      (cudaStream_t)stream);
  });
}

int exa_percentiles_from_hist_values(const uint64_t* hist, const double* values, int nbins, int is_f32,
                                     double q_lo, double q_hi, double* mn, double* mx) {
  return guarded_static([&] {
    return exa::percentiles_from_hist_values(hist, values, nbins, is_f32, q_lo, q_hi, mn, mx);
  });
}

int exa_slab_run(exa_engine* e, const uint16_t* slab_dev, int D, int H, int W,
                 const exa_predict_params* p, int row_begin, int row_end, void* stream) {
  return guarded(e, [&] {
    EXA_CHECK(p != nullptr, "slab_run: null params");
    return e->impl.slab_run(slab_dev, D, H, W, *p, row_begin, row_end, (cudaStream_t)stream);
  });
}

int exa_slab_partial(exa_engine* e, float* halo_dev, void* stream) {
  return guarded(e, [&] { return e->impl.slab_partial(halo_dev, (cudaStream_t)stream); });
}

int exa_slab_stitch(exa_engine* e, const float* seed_dev, float* out_dev, void* stream) {
  return guarded(e, [&] { return e->impl.slab_stitch(seed_dev, out_dev, 0, (cudaStream_t)stream); });
}

int exa_slab_stitch_strided(exa_engine* e, const float* seed_dev, float* out_dev,
                            int64_t channel_stride, void* stream) {
  return guarded(e, [&] {
    return e->impl.slab_stitch(seed_dev, out_dev, channel_stride, (cudaStream_t)stream);
  });
}

int exa_slab_predict(exa_engine* e, const uint16_t* slab_dev, int D, int H, int W,
                     const exa_predict_params* p, int row_begin, int row_end, float* out_dev,
                     int64_t channel_stride, float* out_host, int64_t host_channel_stride,
                     float* halo_dev, void* stream) {
  return guarded(e, [&] {
    EXA_CHECK(p != nullptr, "slab_predict: null params");
    return e->impl.slab_predict(slab_dev, D, H, W, *p, row_begin, row_end, out_dev, channel_stride,
                                out_host, host_channel_stride, halo_dev, (cudaStream_t)stream);
  });
}

int exa_slab_finish(exa_engine* e, const float* seed_dev, float* out_dev, int64_t channel_stride,
                    float* out_host, int64_t host_channel_stride, void* stream) {
  return guarded(e, [&] {
    return e->impl.slab_finish(seed_dev, out_dev, channel_stride, out_host, host_channel_stride,
                               (cudaStream_t)stream);
  });
}

int exa_set_peer_outputs(exa_engine* e, float* local_base, int64_t elems, float* const* peer_bases,
                         int n_peers) {
  return guarded(e, [&] { return e->impl.set_peer_outputs(local_base, elems, peer_bases, n_peers); });
}

int exa_affinities_to_segmentation(int device, const float* aff_host, int D, int H, int W,
                                   const double* thresholds, int n_thresholds, double aff_low,
                                   double aff_high, int64_t min_segment_size, uint64_t* seg_host,
                                   int64_t* n_fragments, int64_t* n_segments) {
  return guarded_static([&] {
    return exa::affinities_to_segmentation_host(device, aff_host, D, H, W, thresholds, n_thresholds,
                                                aff_low, aff_high, min_segment_size, seg_host,
                                                n_fragments, n_segments);
  });
}

int exa_ws_release_memory(void) {
  exa::ws_release_memory();
  return EXA_OK;
}

int exa_ws_last_profile(double* out, int n) {
  if (!out || n < 0) return EXA_ERR_INVALID;
  exa::ws_last_profile(out, n);
  return EXA_OK;
}

int exa_region_agglomerate(int device, uint32_t n_fragments, int64_t n_edges, const uint32_t* eu,
                           const uint32_t* ev, const uint64_t* qsum, const uint32_t* count,
                           double threshold, uint32_t* root_out) {
  return guarded_static([&] {
    return exa::region_agglomerate(device, n_fragments, n_edges, eu, ev, qsum, count, threshold,
                                   root_out);
  });
}

int exa_affinities_to_segmentation_device(const float* aff_dev, int D, int H, int W,
                                          const double* thresholds, int n_thresholds,
                                          double aff_low, double aff_high,
                                          int64_t min_segment_size, uint64_t* seg_dev,
                                          int64_t* n_fragments, int64_t* n_segments, void* stream) {
  return guarded_static([&] {
    return exa::affinities_to_segmentation_device(aff_dev, D, H, W, thresholds, n_thresholds,
                                                  aff_low, aff_high, min_segment_size, seg_dev,
                                                  n_fragments, n_segments, (cudaStream_t)stream);
  });
}

int exa_count_patches(int D, int H, int W, const int32_t patch[3], const int32_t overlap[3]) {
  if (!patch || !overlap) return EXA_ERR_INVALID;
  const int dims[3] = {D, H, W};
  long long n = 1;
  for (int i = 0; i < 3; ++i) {
    if (patch[i] <= 0 || overlap[i] < 0 || overlap[i] >= patch[i]) return EXA_ERR_INVALID;
    n *= (long long)exa::axis_starts(dims[i], patch[i], overlap[i]).size();
  }
  return n > 0x7fffffffLL ? EXA_ERR_INVALID : (int)n;
}

int exa_patch_starts(int D, int H, int W, const int32_t patch[3], const int32_t overlap[3],
                     int32_t* starts, int capacity) {
  const int n = exa_count_patches(D, H, W, patch, overlap);
  if (n < 0) return n;
  if (!starts || capacity < n) return EXA_ERR_INVALID;
  const std::vector<int> sz = exa::axis_starts(D, patch[0], overlap[0]);
  const std::vector<int> sy = exa::axis_starts(H, patch[1], overlap[1]);
  const std::vector<int> sx = exa::axis_starts(W, patch[2], overlap[2]);
  size_t i = 0;
  for (int z : sz)
    for (int y : sy)
      for (int x : sx) {
        starts[i++] = z;
        starts[i++] = y;
        starts[i++] = x;
      }
  return n;
}

int exa_profile_begin(exa_engine* e) {
  return guarded(e, [&] { return e->impl.profile_begin(); });
}

int exa_profile_end(exa_engine* e, double* ms_by_category, int64_t* launches_by_category, int n) {
  return guarded(e, [&] { return e->impl.profile_end(ms_by_category, launches_by_category, n); });
}

int exa_profile_layers(exa_engine* e, double* ms, int64_t* launches, int32_t* kind, int n) {
  return guarded(e, [&] { return e->impl.profile_layers(ms, launches, kind, n); });
}

int64_t exa_launch_count(const exa_engine* e) { return e ? e->impl.launches : -1; }

// ---- training step (trainer.h) --------------------------------------------------------------
int exa_train_create(int device, int precision, exa_trainer** out) {
  if (!out) return EXA_ERR_INVALID;
  *out = nullptr;
  return guarded_static([&] {
    exa_trainer* t = new (std::nothrow) exa_trainer(device, precision);
    if (!t) return exa::Status::Err("out of host memory");
    exa::Status s = t->impl.init();
    if (!s.ok) {
      delete t;
      return s;
    }
    *out = t;
    return exa::Status::OK();
  });
}

int exa_train_destroy(exa_trainer* t) {
  if (!t) return EXA_ERR_INVALID;
  delete t;
  return EXA_OK;
}

const char* exa_train_last_error(const exa_trainer* t) {
  return t ? t->impl.last_error.c_str() : g_create_error.c_str();
}

int exa_train_bind(exa_trainer* t, const char* name, float* dev_ptr, const int64_t* shape,
                   int ndim) {
  return guarded_train(t, [&] { return t->impl.bind(name, dev_ptr, shape, ndim); });
}

int exa_train_grad_elems(exa_trainer* t, int64_t* n) {
  return guarded_train(t, [&] { return t->impl.grad_elems(n); });
}

int exa_train_grad_slot(exa_trainer* t, const char* name, int64_t* offset, int64_t* numel) {
  return guarded_train(t, [&] { return t->impl.grad_slot(name, offset, numel); });
}

int exa_train_out_channels(exa_trainer* t) {
  if (!t) return EXA_ERR_INVALID;
  int64_t n = 0;
  const int code = guarded_train(t, [&] { return t->impl.grad_elems(&n); });  // resolves
  return code < 0 ? code : t->impl.out_channels();
}

int exa_train_forward(exa_trainer* t, const float* x_dev, int batch, const int32_t patch[3],
                      float* logits_dev, void* stream) {
  return guarded_train(t, [&] {
    if (!patch) return exa::Status::Err("train_forward: null patch");
    return t->impl.forward(x_dev, batch, patch, logits_dev, (cudaStream_t)stream);
  });
}

int exa_train_backward(exa_trainer* t, const float* x_dev, const float* grad_logits_dev,
                       float* grads_dev, void* stream) {
  return guarded_train(
      t, [&] { return t->impl.backward(x_dev, grad_logits_dev, grads_dev, (cudaStream_t)stream); });
}

int exa_train_profile_begin(exa_trainer* t) {
  return guarded_train(t, [&] { return t->impl.profile_begin(); });
}

int exa_train_profile_end(exa_trainer* t, double* ms_by_category, int64_t* launches_by_category,
                          int n) {
  return guarded_train(
      t, [&] { return t->impl.profile_end(ms_by_category, launches_by_category, n); });
}

int64_t exa_train_launch_count(const exa_trainer* t) { return t ? t->impl.launches : -1; }
int64_t exa_train_workspace_bytes(const exa_trainer* t) {
  return t ? (int64_t)t->impl.workspace_bytes() : -1;
}

int exa_conv3d_weight_grad(int device, int precision, const void* x_dev, const void* dz_dev, int B,
                           int D, int H, int W, int cin, int cout, float* dw_dev, void* stream) {
  return guarded_static([&] {
    return exa::conv3d_weight_grad(device, precision, x_dev, dz_dev, B, D, H, W, cin, cout, dw_dev,
                                   (cudaStream_t)stream);
  });
}

int exa_conv3d_data_grad(int device, int precision, const void* dz_dev, const float* w_dev, int B,
                         int D, int H, int W, int cin, int cout, void* dx_dev, void* stream) {
  return guarded_static([&] {
    return exa::conv3d_data_grad(device, precision, dz_dev, w_dev, B, D, H, W, cin, cout, dx_dev,
                                 (cudaStream_t)stream);
  });
}

int exa_bce_with_logits(const float* logits_dev, const float* target_dev, int64_t n,
                        float grad_scale, double* loss_sum_dev, float* grad_dev, void* stream) {
  return guarded_static([&] {
    if (n <= 0) return exa::Status::Err("bce_with_logits: n must be positive");
    return exa::launch_bce_with_logits(logits_dev, target_dev, (size_t)n, grad_scale, loss_sum_dev,
                                       grad_dev, (cudaStream_t)stream);
  });
}

}  // extern "C"
