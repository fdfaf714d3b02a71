// K7: affinities -> segmentation on the GPU (reference inference.py:196-237); see watershed.cu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace exa {

// aff: device float32 (3, D, H, W); seg: device uint64 (D, H, W).  Synchronises `s` (round counters
// come back to the host, and the tail of the merge queue runs there).
Status affinities_to_segmentation_device(const float* aff, int D, int H, int W,
                                         const double* thresholds, int n_thresholds, double aff_low,
                                         double aff_high, int64_t min_segment_size, uint64_t* seg,
                                         int64_t* n_fragments, int64_t* n_segments, cudaStream_t s);
// same with host buffers (copies in and out on `device`)
Status affinities_to_segmentation_host(int device, const float* aff, int D, int H, int W,
                                       const double* thresholds, int n_thresholds, double aff_low,
                                       double aff_high, int64_t min_segment_size, uint64_t* seg,
                                       int64_t* n_fragments, int64_t* n_segments);

// wall milliseconds of the phases of the last affinities_to_segmentation call in this process:
// [0] fragments, [1] region graph, [2] parallel agglomeration rounds, [3] host queue, [4] sizes +
// relabel, then counts: [5] parallel rounds, [6] region-graph edges, [7] edges given to the host queue
void ws_last_profile(double* out, int n);

// gives the device memory cached by this file's block cache (current device) back to the driver
void ws_release_memory();

// The agglomeration step alone on a region graph given as host arrays: edges eu[i] < ev[i] (fragment
// ids 1..n_fragments, every pair at most once, sorted by (eu, ev): the index is the tie-break rank),
// qsum[i] = sum of the affinities between the two in 32.32 fixed point, count[i] = faces.
// device < 0: the exact host queue only (no GPU needed); otherwise parallel rounds on that GPU and
// the host queue for the rest, as affinities_to_segmentation_device does.  root_out[0..n_fragments]:
// the smallest fragment id of the region every fragment ends up in.
Status region_agglomerate(int device, uint32_t n_fragments, int64_t n_edges, const uint32_t* eu,
                          const uint32_t* ev, const uint64_t* qsum, const uint32_t* count,
                          double threshold, uint32_t* root_out);

}  // namespace exa
