// Host launcher for the tensor-core stem conv (kernel in conv_stem.cuh) and the input split.
#include "conv_stem.cuh"

#include "kernels.h"
#include "tmap.h"

namespace exa {

Status launch_stem_tc(const __nv_bfloat16* xs, const __nv_bfloat16* w_band, const float* bias,
                      const Act& out, int num_sms, cudaStream_t s) {
  EXA_CHECK(!out.fp32 && out.C == 32 && out.cstride % 16 == 0 && out.coff % 16 == 0 &&
                out.coff + 32 <= out.cstride,
            "stem_tc: output must be a 32-channel bf16 slot aligned to 16 channels");
  EXA_CHECK(out.D % 16 == 0 && out.H % 8 == 0 && out.W % 4 == 0, "stem_tc: patch dims");
  StemTcArgs a{};
  a.B = out.B;
  a.P[0] = out.D; a.P[1] = out.H; a.P[2] = out.W;
  a.ntz = out.D / 16; a.nty = out.H / 8; a.ntx = ceil_div(out.W, 4 * STEM_SUBS);
  a.tiles_total = out.B * a.ntz * a.nty * a.ntx;
  a.bias = bias;
  a.out = (__nv_bfloat16*)out.ptr;
  a.cstride = out.cstride;
  a.coff = out.coff;
  CUtensorMap tx, tw;
  {
    // normalised input as interleaved (hi, lo) bf16 pairs, x innermost, rows padded to W + 8
    // voxels with voxel x at index x + 1 (launch_stem_split): 4-D (2*Xp, Y, Z, B) in bf16
    // elements, box (64 = 32 voxels, 10, 18, 1)
    const uint64_t wp = (uint64_t)out.W + 8;
    uint64_t dims[4] = {2 * wp, (uint64_t)out.H, (uint64_t)out.D, (uint64_t)out.B};
    uint64_t strides[3] = {wp * 4, wp * out.H * 4, wp * out.H * out.D * 4};
    uint32_t box[4] = {64, 10, 18, 1};  // 32 voxels per row: one box serves six x sub-tiles
    EXA_TRY(make_tmap_bf16(&tx, (void*)xs, 4, dims, strides, box, 128));
  }
  {
    // band matrices: 3-D (16 = 8 x' x (hi, lo), 128 = (xo, c), 9 = (kz, ky)), box (16, 128, 1)
    uint64_t dims[3] = {16, 128, 9};
    uint64_t strides[2] = {32, 32 * 128};
    uint32_t box[3] = {16, 128, 1};
    EXA_TRY(make_tmap_bf16(&tw, (void*)w_band, 3, dims, strides, box, 32));
  }
  static bool configured = false;
  if (!configured) {
    EXA_CUDA(cudaFuncSetAttribute(stem_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  StemTcSmem::TOTAL));
    configured = true;
  }
  const int grid = a.tiles_total < num_sms ? a.tiles_total : num_sms;
  stem_tc_kernel<<<grid, ZF_THREADS, StemTcSmem::TOTAL, s>>>(tx, tw, a);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

}  // namespace exa
