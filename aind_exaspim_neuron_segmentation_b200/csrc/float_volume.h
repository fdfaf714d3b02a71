// Floating-point images for the predict path (reference inference.py:79-80); see float_volume.cu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace exa {

// vol_dev: n float32 (is_double = 0) or float64 values on the device.  idx_dev[i] = rank of
// min(vol[i], clip) among the distinct clipped values; table_out (65536 doubles, host) receives those
// values in ascending order, *n_table their number.  Fails when there are more than 65536.
Status compress_float_volume(const void* vol_dev, int is_double, int64_t n, double clip,
                             uint16_t* idx_dev, double* table_out, int* n_table, cudaStream_t s);

// np.percentile(method="linear") from the histogram of the ranks and the value table (host)
Status percentiles_from_hist_values(const uint64_t* hist, const double* values, int nbins, int is_f32,
                                    double q_lo, double q_hi, double* mn, double* mx);

}  // namespace exa
