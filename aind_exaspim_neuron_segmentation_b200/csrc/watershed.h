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

// the host merge queue alone (no GPU): edges (a << 32 | b, a < b) with summed affinity and face
// count -> root fragment of every fragment 0..n_fragments (root_out has n_fragments + 1 entries)
Status region_agglomerate(uint32_t n_fragments, int64_t n_edges, const uint64_t* pair_keys,
                          const double* sums, const int32_t* counts, double threshold,
                          uint32_t* root_out);

}  // namespace exa
