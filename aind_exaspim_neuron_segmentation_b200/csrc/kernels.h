// Internal launch API of the sm_100a kernels (host side).  Everything is
// stream-ordered and returns a Status; nothing here synchronises.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace exa {

// NDHWC activation view.  `C` channels are addressed inside a buffer whose voxel
// stride is `cstride` elements, starting at channel `coff` (concat buffers).
struct Act {
  void* ptr = nullptr;
  int B = 0, D = 0, H = 0, W = 0;
  int C = 0, cstride = 0, coff = 0;
  bool fp32 = false;
  size_t voxels() const { return (size_t)B * D * H * W; }
};

// Folded weights of the Cin=1 stem conv (reference inc.double_conv.0 + BN .1).
struct StemWeights {
  float w[27][32];
  float b[32];
};

struct PatchSource {
  // (a) patches gathered from a uint16 volume slab with clip + LUT normalisation
  //     (reference inference.py:79-80,188-191; img_util.py:378-379,424-428,526-531)
  const uint16_t* vol = nullptr;
  int gD = 0, gH = 0, gW = 0;  // global volume dims
  int vz0 = 0, vD = 0;         // planes [vz0, vz0+vD) are resident in `vol`
  const float* lut = nullptr;  // [clip+1] float32 normalised values
  int clip = 0;
  const int* starts = nullptr;  // device [B][3] (z, y, x)
  // (b) ready-made float32 patches (B,1,P,P,P)  (operator-level forward)
  const float* x = nullptr;
};

struct HeadParams {
  const float* w = nullptr;  // device [C][c0] (c0 = channels of the last conv; 32 for the fused head)
  const float* b = nullptr;  // device [C]
  const float* w_host = nullptr;  // host copies (passed to K1z as kernel parameters)
  const float* b_host = nullptr;
  float* out = nullptr;      // [B][C][D-2t][H-2t][W-2t]
  int C = 0;
  int trim = 0;
  int apply_sigmoid = 0;
};

// Sliding-window geometry along one axis (reference inference.py:389-393, 101-105).
struct AxisGeom {
  int dim = 0, patch = 0, stride = 0, trim = 0, n = 0;
};

#define EXA_MAX_PEERS 15
struct StitchArgs {
  const float* probs = nullptr;  // [n_slots][C][Pt][Pt][Pt]
  int C = 0;
  AxisGeom az, ay, ax;
  int row_begin = 0, row_end = 0;  // z rows whose patches are resident in `probs`
  int z_begin = 0, z_end = 0;      // output planes to produce
  int y_begin = 0, y_end = 0;      // rows of those planes to produce (0, 0 = all)
  float* out = nullptr;            // [C][z_end-z_begin][H][W] (plane z_begin first), stride below
  size_t out_cstride = 0;          // elements between channels of `out`
  int finalize = 1;                // 1: divide by coverage count; 0: raw partial sums
  const float* seed = nullptr;     // optional partial sums [C][seed_z1-seed_z0][H][W] added FIRST
  int seed_z0 = 0, seed_z1 = 0;
  // fused all-gather (multi-GPU): every finished element is also stored, at the
  // same offset, into the peers' copies of the output array over NVLink (peer-mapped pointers)
  int n_peers = 0;
  float* peer_out[EXA_MAX_PEERS] = {};
};

Status launch_stem(const PatchSource& src, const StemWeights& w, const Act& out, cudaStream_t s);
// tensor-core stem (bf16 mode): input split into bf16 hi/lo, then the Toeplitz-form conv
// xs: interleaved (hi, lo) bf16 pairs, [B][Pz][Py][Px + 8] voxels (2 bf16 each)
Status launch_stem_split(const PatchSource& src, int B, int Pz, int Py, int Px, __nv_bfloat16* xs,
                         cudaStream_t s);
Status launch_stem_tc(const __nv_bfloat16* xs, const __nv_bfloat16* w_band, const float* bias,
                      const Act& out, int num_sms, cudaStream_t s);
Status launch_conv_umma(const Act& in, const Act& out, const __nv_bfloat16* w_packed,
                        const float* bias, const HeadParams* head, int num_sms, cudaStream_t s);
// Output sub-box [lo, hi) (z, y, x) a conv has to produce; voxels outside are left untouched.
struct ConvRegion {
  int lo[3], hi[3];
};
// pair: the cta_group::2 variant (conv_zfold2.cuh), which also takes Cin = 128
bool conv_zfold_supported(const Act& in, int cout, bool pair);
Status launch_conv_zfold(const Act& in, const Act& out, const __nv_bfloat16* w_zfold,
                         const float* bias, const HeadParams* head, const ConvRegion* region,
                         const Act* pool_out, int num_sms, bool pair, cudaStream_t s);
Status launch_conv_fp32(const Act& in, const Act& out, const float* w_packed, const float* bias,
                        cudaStream_t s);
// 1x1x1 head + sigmoid + trim as a kernel of its own (fp32 mode; bf16 models wider than 32)
Status launch_head(const Act& in, const HeadParams& head, cudaStream_t s);
// ConvTranspose3d(k=2, s=2) into a concat slot; w: [8 taps][cin][cout] float32, bias [cout]
Status launch_upconv(const Act& in, const Act& out, const float* w, const float* bias,
                     const ConvRegion* region, cudaStream_t s);
Status launch_maxpool(const Act& in, const Act& out, cudaStream_t s);
// region (optional): only output voxels inside the box are produced
Status launch_upsample(const Act& in, const Act& out, const ConvRegion* region, cudaStream_t s);
Status launch_histogram(const uint16_t* vol, size_t n, int clip, unsigned long long* hist,
                        cudaStream_t s);
Status launch_stitch(const StitchArgs& a, cudaStream_t s);

}  // namespace exa
