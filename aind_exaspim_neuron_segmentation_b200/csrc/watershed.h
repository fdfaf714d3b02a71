// K7: affinities -> segmentation on the GPU (reference inference.py:196-237); see watershed.cu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace exa {

// aff: device float32 (3, D, H, W); seg: device uint64 (D, H, W).  Synchronises `s` (the merge
// queue over the region graph runs on the host).
Status affinities_to_segmentation_device(const float* aff, int D, int H, int W,
                                         const double* thresholds, int n_thresholds, double aff_low,
                                         double aff_high, int64_t min_segment_size, uint64_t* seg,
                                         int64_t* n_fragments, int64_t* n_segments, cudaStream_t s);
// same with host buffers (copies in and out on `device`)
Status affinities_to_segmentation_host(int device, const float* aff, int D, int H, int W,
                                       const double* thresholds, int n_thresholds, double aff_low,
                                       double aff_high, int64_t min_segment_size, uint64_t* seg,
                                       int64_t* n_fragments, int64_t* n_segments);

}  // namespace exa
