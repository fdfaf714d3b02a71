/*
 * exaspim_b200 -- C ABI of the B200-native affinity-prediction hot path.
 *
 * This is the drop-in boundary for ONE path of
 * AllenNeuralDynamics/aind-exaspim-neuron-segmentation:
 *     inference.predict(img, model, affinity_mode=True, patch_shape=(96,96,96))
 * (reference src/aind_exaspim_neuron_segmentation/inference.py:29-126) and the model it
 * drives (machine_learning/unet3d.py:16-336, loaded by inference.py:400-424).
 *
 * Widened, one row of SURVEY.md section 8f at a time, to the callers either side of that path:
 * affinities_to_segmentation (inference.py:196-237; exa_affinities_to_segmentation*), streaming
 * and float inputs, the model variants of unet3d.py:37, and the Trainer's training step
 * (machine_learning/train.py:123-157, 200-223; exa_train_*).
 *
 * The reference is pure Python and has no FFI of its own; its boundary for this path is
 * two Python functions plus the nn.Module call convention.  The
 * entry points below are what a ctypes binding of those functions needs -- see
 * INTEGRATION.md for the reference-side stub.  Each one cites the reference interface it
 * replaces.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a
 * negative code on failure and never throws; exa_last_error() returns a message for the
 * last failure on that engine.  An engine is bound to one CUDA device, owns its packed
 * weights / workspaces, and is not thread-safe.  There is NO CPU fallback: without a CUDA
 * device exa_create fails.
 */
#ifndef EXASPIM_B200_H_
#define EXASPIM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct exa_engine exa_engine;

enum { EXA_PRECISION_BF16 = 0, EXA_PRECISION_FP32 = 1 };
enum { EXA_DTYPE_F32 = 0, EXA_DTYPE_I64 = 1 };

enum {
  EXA_OK = 0,
  EXA_ERR_INVALID = -1, /* bad argument / unsupported configuration */
  EXA_ERR_CUDA = -2,    /* CUDA runtime or driver failure           */
  EXA_ERR_STATE = -3    /* call sequence error (e.g. weights not finalised) */
};

/* Keyword arguments of reference predict(), inference.py:29-40.  batch is the number of
 * patches the engine runs per wave (numerically irrelevant, SURVEY.md 8a invariants);
 * 0 selects a default. */
typedef struct {
  int32_t patch[3];    /* patch_shape (z, y, x); multiples of 16 */
  int32_t overlap[3];  /* overlap (z, y, x)                      */
  int32_t trim;        /* trim                                   */
  int32_t brightness_clip;
  double pct_lo, pct_hi; /* normalization_percentiles            */
  int32_t batch;
  int32_t reserved;
} exa_predict_params;

/* ---- engine lifetime ------------------------------------------------------------ */
int exa_create(int device, int precision, exa_engine** out);
int exa_destroy(exa_engine* e);
const char* exa_last_error(const exa_engine* e); /* e may be NULL: last create error */
const char* exa_version(void);

/* ---- weights: replaces UNet3D.load_state_dict, inference.py:420-421 --------------
 * One call per state_dict entry (128 entries for the reference architecture, names and
 * shapes as in unet3d.py:64-75,142-149,318).  Data is copied from host memory.
 * exa_finalize_weights checks the key set strictly, folds eval-mode BatchNorm
 * (unet3d.py:144,147) into the conv weights in fp32 and packs them for the kernels. */
int exa_load_weight(exa_engine* e, const char* name, const void* data, const int64_t* shape,
                    int ndim, int dtype);
int exa_finalize_weights(exa_engine* e);
int exa_out_channels(const exa_engine* e); /* 3 (affinity_mode) or 1 */

/* ---- operator level: replaces UNet3D.forward, unet3d.py:77-105 --------------------
 * x: device float32 (B,1,Pz,Py,Px); logits: device float32 (B,C,Pz,Py,Px).
 * stream: cudaStream_t (may be NULL).  Asynchronous. */
int exa_forward(exa_engine* e, const float* x, float* logits, int batch, const int32_t patch[3],
                void* stream);

/* ---- whole path, host buffers: replaces predict(), inference.py:29-126 -------------
 * vol: host uint16 (D,H,W), not modified.  out: host float32 (C,D,H,W), fully written
 * (uncovered border voxels are 0.0 exactly as in the reference).  Synchronous. */
int exa_predict(exa_engine* e, const uint16_t* vol, int D, int H, int W,
                const exa_predict_params* p, float* out);

/* Progress reporting for the tqdm bar of predict(verbose=True) (inference.py:94-95,118-120):
 * cb(user, patches_done, patches_total) is called from a CUDA host-function thread each time a
 * wave of patches has actually FINISHED on the device (cudaLaunchHostFunc on the engine's
 * stream), while exa_predict is still running.  cb must not call into this library or CUDA.
 * NULL switches it off (default: no host functions are enqueued at all). */
typedef void (*exa_progress_fn)(void* user, int64_t patches_done, int64_t patches_total);
int exa_set_progress_callback(exa_engine* e, exa_progress_fn cb, void* user);

/* Same with device-resident buffers (the bench's device-timed region). Asynchronous on
 * `stream` except for one small D2H of the 1001-bin histogram. */
int exa_predict_device(exa_engine* e, const uint16_t* vol_dev, int D, int H, int W,
                       const exa_predict_params* p, float* out_dev, void* stream);

/* ---- slab pieces for z-row sharding across GPUs --------------------
 * A volume's patch grid has nz rows of patches along z.  A rank owns rows
 * [row_begin,row_end); it needs input planes [exa_slab_in_z0, exa_slab_in_z1) resident and
 * produces output planes [exa_slab_out_z0, exa_slab_out_z1).  The planes
 * [halo_z0, halo_z1) at its upper end are also covered by the next rank's first row: the
 * rank sends raw partial sums for them (exa_slab_partial) and the next rank seeds its
 * stitch with them so that the summation order matches inference.py:99-116 bit for bit. */
typedef struct {
  int32_t nz, ny, nx;              /* patch grid                               */
  int32_t n_patches;               /* nz*ny*nx  (count_patches, inference.py:340-365) */
  int32_t in_z0, in_z1;            /* input planes needed                      */
  int32_t out_z0, out_z1;          /* output planes owned                      */
  int32_t halo_z0, halo_z1;        /* planes to hand to the next rank (empty if last) */
  int32_t seed_z0, seed_z1;        /* planes to receive from the previous rank */
} exa_slab_plan;

int exa_plan_slab(int D, int H, int W, const exa_predict_params* p, int row_begin, int row_end,
                  exa_slab_plan* plan);
/* histogram of min(v, clip) over n uint16 values on the device -> hist_dev[clip+1] (uint64) */
int exa_histogram(exa_engine* e, const uint16_t* vol_dev, int64_t n, int clip,
                  uint64_t* hist_dev, void* stream);
/* exact np.percentile(method="linear") from a histogram (host, no GPU needed):
 * inference.py:80 -> img_util.py:526 */
int exa_percentiles_from_hist(const uint64_t* hist, int nbins, double q_lo, double q_hi,
                              double* mn, double* mx);
/* normalisation scalars (img_util.py:526-531) for the following slab calls */
int exa_set_normalization(exa_engine* e, double mn, double mx, int clip);
/* Floating-point images (inference.py:79-80 accepts any dtype).  exa_compress_float_volume turns n
 * float32 (is_double = 0) / float64 values on the device into uint16 ranks of min(x, clip) among its
 * distinct values, which it returns in ascending order in table_out (65536 doubles, host); more than
 * 65536 distinct clipped values, or NaN, is an error.  exa_percentiles_from_hist_values is
 * np.percentile(method="linear") from the histogram of the ranks (exa_histogram) and that table
 * (is_f32: the image was float32 -- numpy takes the neighbour difference in the array's dtype);
 * exa_set_normalization_table installs the lookup table of the normalised values
 * (img_util.py:527-531 in float64) for the following slab calls, whose brightness_clip is n - 1. */
int exa_compress_float_volume(const void* vol_dev, int is_double, int64_t n, double clip,
                              uint16_t* idx_dev, double* table_out, int* n_table, void* stream);
int exa_percentiles_from_hist_values(const uint64_t* hist, const double* values, int nbins, int is_f32,
                                     double q_lo, double q_hi, double* mn, double* mx);
int exa_set_normalization_table(exa_engine* e, const double* values, int n, double mn, double mx);
/* run all patches of rows [row_begin,row_end); slab_dev holds planes [in_z0,in_z1) */
int exa_slab_run(exa_engine* e, const uint16_t* slab_dev, int D, int H, int W,
                 const exa_predict_params* p, int row_begin, int row_end, void* stream);
/* raw partial sums of the halo planes -> halo_dev (C, halo_z1-halo_z0, H, W) */
int exa_slab_partial(exa_engine* e, float* halo_dev, void* stream);
/* finished planes [out_z0,out_z1) -> out_dev (C, out_z1-out_z0, H, W); seed_dev may be NULL */
int exa_slab_stitch(exa_engine* e, const float* seed_dev, float* out_dev, void* stream);
/* same, writing straight into a larger channel-major array: out_dev points at plane out_z0 of
 * channel 0 and consecutive channels are channel_stride elements apart (e.g. D*H*W of the full
 * (C, D, H, W) output, so that the gather of the other ranks' planes needs no re-packing) */
int exa_slab_stitch_strided(exa_engine* e, const float* seed_dev, float* out_dev,
                            int64_t channel_stride, void* stream);
/* Rows [row_begin,row_end) as one pipelined job (the sharded form of exa_predict's inner loop,
 * inference.py:93-125): the rows run in groups, every finished plane is stitched into out_dev
 * (plane out_z0 first, channels channel_stride elements apart) and, when out_host is not NULL,
 * copied to out_host (pinned memory recommended; same layout with host_channel_stride) on a second
 * stream while the next group computes.  Needs exa_set_normalization first.  The partial sums for
 * the next rank go to halo_dev (as exa_slab_partial; NULL only for the last rows).  The planes
 * shared with the previous rank [seed_z0,seed_z1) are NOT produced here: exchange halo/seed
 * between the ranks, then call exa_slab_finish with the received seed (NULL for the first rows),
 * which stitches and copies those planes and returns once out_host is complete. */
int exa_slab_predict(exa_engine* e, const uint16_t* slab_dev, int D, int H, int W,
                     const exa_predict_params* p, int row_begin, int row_end, float* out_dev,
                     int64_t channel_stride, float* out_host, int64_t host_channel_stride,
                     float* halo_dev, void* stream);
int exa_slab_finish(exa_engine* e, const float* seed_dev, float* out_dev, int64_t channel_stride,
                    float* out_host, int64_t host_channel_stride, void* stream);
/* Fused all-gather: local_base is this GPU's copy of the full (C, D, H, W)
 * output (elems floats) and peer_bases[i] the peer-mapped addresses of the other ranks' copies
 * (CUDA IPC / symmetric memory; at most 15).  While set, the stitch kernel stores every finished
 * element of a destination inside local_base's range to all copies, so no separate gather pass
 * runs; the caller orders a cross-rank barrier after the last stitch.  n_peers = 0 clears it. */
int exa_set_peer_outputs(exa_engine* e, float* local_base, int64_t elems, float* const* peer_bases,
                         int n_peers);

/* ---- downstream of the path: affinities -> segmentation ---------------
 * Replaces affinities_to_segmentation (inference.py:196-237): waterz.agglomerate(affinities,
 * thresholds, aff_threshold_low, aff_threshold_high) -- watershed fragments, region graph,
 * hierarchical merging by 1 - mean affinity up to the last threshold -- followed by
 * remove_small_segments (img_util.py:536-559: segments with more than min_segment_size voxels
 * are kept and renumbered from 1 in order of first appearance).  aff: float32 (3, D, H, W),
 * aff[c][z,y,x] = edge between voxel (z,y,x) and its PREVIOUS neighbour along axis c (waterz's
 * reading of the array); seg: uint64 (D, H, W).  Fragments (with waterz's breadth-first plateau
 * division), the region graph, the parallel agglomeration rounds and the relabelling run on the
 * GPU; the tail of the merge queue runs on the host.  n_fragments / n_segments may be NULL. */
int exa_affinities_to_segmentation(int device, const float* aff_host, int D, int H, int W,
                                   const double* thresholds, int n_thresholds, double aff_low,
                                   double aff_high, int64_t min_segment_size, uint64_t* seg_host,
                                   int64_t* n_fragments, int64_t* n_segments);
/* same on device buffers of the current device (e.g. the output of exa_predict_device) */
int exa_affinities_to_segmentation_device(const float* aff_dev, int D, int H, int W,
                                          const double* thresholds, int n_thresholds,
                                          double aff_low, double aff_high,
                                          int64_t min_segment_size, uint64_t* seg_dev,
                                          int64_t* n_fragments, int64_t* n_segments, void* stream);

/* exa_affinities_to_segmentation* keep the device blocks they used in a per-device cache (the next
 * call of the same size allocates nothing); this gives them back to the driver (current device). */
int exa_ws_release_memory(void);

/* measurement (no reference analogue): wall milliseconds of the phases of the last
 * exa_affinities_to_segmentation[_device] call of this process -- [0] fragments, [1] region graph,
 * [2] parallel agglomeration rounds, [3] host queue, [4] sizes + relabel -- then counts:
 * [5] parallel rounds, [6] region-graph edges, [7] edges handed to the host queue.  n <= 8. */
int exa_ws_last_profile(double* out, int n);

/* the agglomeration step of the function above on its own, on a region graph in host arrays:
 * edges eu[i] < ev[i] (fragment ids 1..n_fragments, each pair once, sorted by (eu, ev) -- the index
 * is the tie-break rank), qsum[i] = sum of the affinities on the faces between the two fragments in
 * 32.32 fixed point (llrint(clamp(a, 0, 1) * 2^32) per face), count[i] = number of faces.
 * device < 0: the exact host merge queue only (no GPU needed); device >= 0: parallel rounds on that
 * GPU, host queue for the rest.  root_out[0..n_fragments] receives the smallest fragment id of the
 * region every fragment is merged into while the smallest 1 - mean affinity is < threshold. */
int exa_region_agglomerate(int device, uint32_t n_fragments, int64_t n_edges, const uint32_t* eu,
                           const uint32_t* ev, const uint64_t* qsum, const uint32_t* count,
                           double threshold, uint32_t* root_out);

/* host helpers mirroring count_patches / generate_patch_starts (inference.py:340-397);
 * starts receives n_patches*3 int32 (z,y,x) in the reference's order */
int exa_count_patches(int D, int H, int W, const int32_t patch[3], const int32_t overlap[3]);
int exa_patch_starts(int D, int H, int W, const int32_t patch[3], const int32_t overlap[3],
                     int32_t* starts, int capacity);

/* ---- measurement (bench.py roofline; no reference analogue) ------------------------
 * Between exa_profile_begin and exa_profile_end every kernel launch of this engine is
 * bracketed by two CUDA events on its stream; exa_profile_end synchronises and returns
 * summed device milliseconds and launch counts per category:
 * 0 histogram, 1 stem(gather+normalise+inc.0), 2 conv3x3x3, 3 max-pool, 4 upsample,
 * 5 head (fp32 mode only; fused into the last conv otherwise), 6 stitch.  n >= 7. */
enum { EXA_PROFILE_CATEGORIES = 7 };
int exa_profile_begin(exa_engine* e);
int exa_profile_end(exa_engine* e, double* ms_by_category, int64_t* launches_by_category, int n);

/* Per-layer view of the last exa_profile_end: summed device ms and launches of the 3x3x3 conv
 * of layer i (i = 1..17 in the order of unet3d.py:64-74; 0 = the stem, reported under category 1)
 * and the kernel that ran it: 0 none, 1 K1 (per-tap implicit GEMM), 2 K1z (z-folded, one CTA per
 * MMA), 3 K1z2 (z-folded, CTA pairs).  n >= 18. */
int exa_profile_layers(exa_engine* e, double* ms, int64_t* launches, int32_t* kind, int n);

/* number of kernel launches issued by this engine since creation (bench bookkeeping) */
int64_t exa_launch_count(const exa_engine* e);

/* ---- training step (reference machine_learning/train.py:123-157, 200-223) ----
 * What `model.train(); hat_y = model(x); loss = criterion(hat_y, y); loss.backward()` does for
 * the U-Net of unet3d.py:16-105 (trilinear=True, width_multiplier=1, what Trainer builds at
 * train.py:77): the forward pass with BatchNorm3d in TRAINING mode (batch statistics; running
 * statistics updated in place, momentum 0.1, unet3d.py:144,147) and the backward pass to every
 * parameter gradient.  The reference has no FFI; its boundary is nn.Module.__call__ plus
 * autograd, which INTEGRATION.md maps onto these calls through a torch.autograd.Function.
 *
 * Parameters stay where the optimiser keeps them: exa_train_bind() takes the DEVICE pointer of
 * every float32 state_dict entry (the 128 entries of SURVEY.md 8a-1 minus the 18 int64
 * num_batches_tracked counters, which the caller increments).  Every forward reads the current
 * values.  The gradients of one backward are written (not accumulated) into one flat float32
 * device buffer of exa_train_grad_elems() elements; exa_train_grad_slot() gives the slot of a
 * parameter.  A backward differentiates the most recent forward of the same trainer. */
typedef struct exa_trainer exa_trainer;
int exa_train_create(int device, int precision, exa_trainer** out);
int exa_train_destroy(exa_trainer* t);
const char* exa_train_last_error(const exa_trainer* t); /* t may be NULL: last create error */
int exa_train_bind(exa_trainer* t, const char* name, float* dev_ptr, const int64_t* shape,
                   int ndim);
int exa_train_grad_elems(exa_trainer* t, int64_t* n);
int exa_train_grad_slot(exa_trainer* t, const char* name, int64_t* offset, int64_t* numel);
int exa_train_out_channels(exa_trainer* t);
/* x: (B,1,D,H,W) float32 on the device, patch = (D,H,W) multiples of 16;
 * logits: (B,C,D,H,W) float32 on the device (unet3d.py:77-105 in train() mode) */
int exa_train_forward(exa_trainer* t, const float* x_dev, int batch, const int32_t patch[3],
                      float* logits_dev, void* stream);
/* x: the input of that forward; grad_logits: dLoss/dlogits, (B,C,D,H,W) float32 */
int exa_train_backward(exa_trainer* t, const float* x_dev, const float* grad_logits_dev,
                       float* grads_dev, void* stream);
int64_t exa_train_launch_count(const exa_trainer* t);
/* like exa_profile_begin / exa_profile_end: summed device milliseconds and launches per category:
 * 0 weight packing, 1 forward convolutions (tcgen05), 2 BatchNorm forward (statistics + apply),
 * 3 max-pool / upsample / head forward, 4 BatchNorm + LeakyReLU backward, 5 weight gradients
 * (tcgen05), 6 data gradients (tcgen05), 7 head backward, 8 weight gradient of the Cin = 1 stem,
 * 9 split-K reductions of the weight gradients, 10 upsample backward, 11 max-pool backward.
 * n >= 12. */
enum { EXA_TRAIN_PROFILE_CATEGORIES = 12 };
int exa_train_profile_begin(exa_trainer* t);
int exa_train_profile_end(exa_trainer* t, double* ms_by_category, int64_t* launches_by_category,
                          int n);
int64_t exa_train_workspace_bytes(const exa_trainer* t);
/* Operator-level entry points of the backward pass's two tensor-core pieces, for parity tests:
 * what torch.nn.grad.conv3d_weight / conv3d_input compute for the Conv3d(k=3, padding=1) of
 * unet3d.py:143,146.  x / dz / dx: NDHWC device arrays of the precision's element type (bf16
 * or float32; for cin == 1, x is the raw float32 (B,1,D,H,W) input); w / dw: float32 device
 * arrays in the reference's (Cout, Cin, 3, 3, 3) layout.  Synchronous. */
int exa_conv3d_weight_grad(int device, int precision, const void* x_dev, const void* dz_dev, int B,
                           int D, int H, int W, int cin, int cout, float* dw_dev, void* stream);
int exa_conv3d_data_grad(int device, int precision, const void* dz_dev, const float* w_dev, int B,
                         int D, int H, int W, int cin, int cout, void* dx_dev, void* stream);
/* nn.BCEWithLogitsLoss() (train.py:76,222), mean reduction: *loss_sum_dev (double, zeroed by the
 * caller) += sum of the element losses; when grad_dev is not NULL it receives
 * grad_scale * (sigmoid(logit) - target) / n  (grad_scale = the GradScaler factor, train.py:140) */
int exa_bce_with_logits(const float* logits_dev, const float* target_dev, int64_t n,
                        float grad_scale, double* loss_sum_dev, float* grad_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EXASPIM_B200_H_ */
