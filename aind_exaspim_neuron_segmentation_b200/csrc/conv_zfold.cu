// Host launcher for the z-folded halo-tile tcgen05 conv (kernel in conv_zfold.cuh).
#include "conv_zfold.cuh"
#include "conv_zfold2.cuh"

#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"
#include "tmap.h"

namespace exa {

namespace {

template <int CIN, int EPI>
Status launch_zf(const CUtensorMap& tx, const CUtensorMap& tw, const ZfArgs& a, int grid,
                 cudaStream_t s) {
  using S = ZfSmem<CIN>;
  static bool configured = false;
  if (!configured) {
    EXA_CUDA(cudaFuncSetAttribute(conv3x3_zfold_kernel<CIN, EPI>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  conv3x3_zfold_kernel<CIN, EPI><<<grid, ZF_THREADS, S::TOTAL, s>>>(tx, tw, a);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

template <int CIN, int EPI>
Status launch_zf2(const CUtensorMap& tx, const CUtensorMap& tw, const ZfArgs& a, int grid,
                  cudaStream_t s) {
  using S = Zf2Smem<CIN>;
  static bool configured = false;
  if (!configured) {
    EXA_CUDA(cudaFuncSetAttribute(conv3x3_zfold2_kernel<CIN, EPI>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  conv3x3_zfold2_kernel<CIN, EPI><<<grid, ZF_THREADS, S::TOTAL, s>>>(tx, tw, a);  // clusters of 2
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

}  // namespace

bool conv_zfold_supported(const Act& in, int cout, bool pair) {
  const bool cin_ok = in.C == 32 || in.C == 64 || (pair && in.C == 128);
  // pair mode tolerates a ragged last tile row (H % 16 != 0: out-of-range rows are zero-filled by
  // TMA and masked in the epilogue), which admits the 24^3 level (up2.conv.double_conv.3)
  return !in.fp32 && cin_ok && (cout == 32 || cout == 64) && in.W % 8 == 0 &&
         (in.H % 16 == 0 || (pair && in.H % 8 == 0)) && in.D >= 2;
}

// w_zfold: bf16 [9 taps (ky,kx)][3 (kz = 2,1,0)][Cout][Cin]
Status launch_conv_zfold(const Act& in, const Act& out, const __nv_bfloat16* w_zfold,
                         const float* bias, const HeadParams* head, const ConvRegion* region,
                         const Act* pool_out, int num_sms, bool pair, cudaStream_t s) {
  const int Cin = in.C, Cout = head ? 32 : out.C;
  EXA_CHECK(conv_zfold_supported(in, Cout, pair), "conv_zfold: unsupported layer shape");
  ZfArgs a{};
  a.B = in.B; a.D = in.D; a.H = in.H; a.W = in.W;
  int x0 = 0, x1 = in.W, y0 = 0, y1 = in.H, z0 = 0, z1 = in.D;
  if (region) {
    x0 = region->lo[2]; x1 = region->hi[2];
    y0 = region->lo[1]; y1 = region->hi[1];
    z0 = region->lo[0]; z1 = region->hi[0];
    EXA_CHECK(0 <= x0 && x0 < x1 && x1 <= in.W && 0 <= y0 && y0 < y1 && y1 <= in.H && 0 <= z0 &&
                  z0 < z1 && z1 <= in.D,
              "conv_zfold: bad output region");
  }
  a.ox = x0; a.oy = y0; a.oz = z0;
  a.ntx = ceil_div(x1 - x0, 8);
  a.nty = ceil_div(y1 - y0, 16);
  a.nzp = z1 - z0;
  a.n_halves = Cout / 32;
  a.tiles_total = in.B * a.nty * a.ntx;
  a.bias = bias;
  if (head) {
    EXA_CHECK(Cout == 32, "fused head needs Cout == 32");
    EXA_CHECK(head->C >= 1 && head->C <= 8 && head->w_host && head->b_host,
              "conv_zfold: fused head needs 1..8 channels and host weights");
    for (int oc = 0; oc < head->C; ++oc) {
      for (int j = 0; j < 32; ++j) a.head_w[oc][j] = head->w_host[oc * 32 + j];
      a.head_b[oc] = head->b_host[oc];
    }
    a.head_out = head->out;
    a.head_c = head->C; a.trim = head->trim; a.apply_sigmoid = head->apply_sigmoid;
  } else {
    EXA_CHECK(!out.fp32 && out.B == in.B && out.D == in.D && out.H == in.H && out.W == in.W,
              "conv_zfold: output shape mismatch");
    EXA_CHECK((out.cstride % 8) == 0 && (out.coff % 8) == 0, "conv_zfold: output alignment");
    a.out = (__nv_bfloat16*)out.ptr; a.out_cstride = out.cstride; a.out_coff = out.coff;
    if (pool_out) {
      EXA_CHECK(!region, "conv_zfold: fused pool needs the full output region");
      EXA_CHECK(in.D % 2 == 0 && pool_out->D * 2 == in.D && pool_out->H * 2 == in.H &&
                    pool_out->W * 2 == in.W && pool_out->B == in.B && pool_out->C == Cout &&
                    !pool_out->fp32 && pool_out->cstride % 8 == 0 && pool_out->coff % 8 == 0,
                "conv_zfold: fused pool output shape mismatch");
      a.pool_out = (__nv_bfloat16*)pool_out->ptr;
      a.pool_cstride = pool_out->cstride;
      a.pool_coff = pool_out->coff;
    }
  }

  CUtensorMap tx, tw;
  // pair mode (conv_zfold2.cuh): channel boxes of at most 64 (128 B swizzle span), weight rows in
  // chunks of 16 so that each CTA of a pair can fetch its 48 of the 96 [kz][cout] rows
  const int cbox = pair && Cin > 64 ? 64 : Cin;
  {
    // activations: 5-D (C, W, H, D, B), box = (cbox, 10, 18, 1, 1): the halo tile of one plane
    const uint64_t cs = (uint64_t)in.cstride * 2;
    uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)in.W, (uint64_t)in.H, (uint64_t)in.D,
                        (uint64_t)in.B};
    uint64_t strides[4] = {cs, cs * in.W, cs * in.W * in.H, cs * in.W * in.H * in.D};
    uint32_t box[5] = {(uint32_t)cbox, 10, 18, 1, 1};
    void* base = (void*)((__nv_bfloat16*)in.ptr + in.coff);
    EXA_TRY(make_tmap_bf16(&tx, base, 5, dims, strides, box, cbox * 2));
  }
  {
    // weights: 4-D (Cin, Cout, 3, 9); box = (Cin, 32, 3, 1): one tap, one 32-channel slice;
    // pair mode: box = (cbox, 16, 1, 1)
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)Cout, 3, 9};
    uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)Cin * Cout * 2,
                           (uint64_t)Cin * Cout * 3 * 2};
    uint32_t box[4] = {(uint32_t)cbox, pair ? 16u : 32u, pair ? 1u : 3u, 1};
    EXA_TRY(make_tmap_bf16(&tw, (void*)w_zfold, 4, dims, strides, box, cbox * 2));
  }
  int grid;
  if (pair) {
    grid = num_sms - num_sms % (2 * a.n_halves);
    const int rounds = (a.tiles_total + 1) / 2;
    if (grid > 2 * rounds * a.n_halves) grid = 2 * rounds * a.n_halves;
  } else {
    grid = num_sms - num_sms % a.n_halves;
    if (grid > a.tiles_total * a.n_halves) grid = a.tiles_total * a.n_halves;
  }
  {
    static const char* dbg_env = getenv("EXA_ZF_DBG");  // 8: print issuer cycles/plane and SM clock
    a.dbg = dbg_env ? (atoi(dbg_env) & 8) : 0;
  }
  static long long* dbg_dev = nullptr;
  if (a.dbg & 8) {
    if (!dbg_dev) EXA_CUDA(cudaMalloc(&dbg_dev, sizeof(long long) * 4 * 1024));
    EXA_CUDA(cudaMemsetAsync(dbg_dev, 0, sizeof(long long) * 4 * 1024, s));
    a.dbg_out = dbg_dev;
  }
  struct DbgPrint {  // prints after the launch below (synchronises: development only)
    bool on; int grid, cin, cout, head; cudaStream_t s; long long* dev;
    ~DbgPrint() {
      if (!on) return;
      cudaStreamSynchronize(s);
      static long long h[4 * 1024];
      cudaMemcpy(h, dev, sizeof(long long) * 4 * grid, cudaMemcpyDeviceToHost);
      double cyc = 0, ns = 0, pl = 0, mx = 0;
      for (int i = 0; i < grid; ++i) {
        cyc += (double)h[i * 4]; ns += (double)h[i * 4 + 1]; pl += (double)h[i * 4 + 2];
        if ((double)h[i * 4] > mx) mx = (double)h[i * 4];
      }
      fprintf(stderr, "[zfold cin=%d cout=%d head=%d grid=%d] issuer: %.0f cycles/plane, %.0f planes/cta, "
              "SM clock %.0f MHz, longest cta %.3f ms\n", cin, cout, head, grid, cyc / pl, pl / grid,
              1e3 * cyc / ns, mx / (1e3 * cyc / ns) * 1e-3);
    }
  } dbg_print{(a.dbg & 8) != 0, grid, Cin, Cout, head != nullptr, s, dbg_dev};
  if (pair) {
    if (Cin == 32) {
      if (head) return launch_zf2<32, EPI_HEAD>(tx, tw, a, grid, s);
      return launch_zf2<32, EPI_STORE>(tx, tw, a, grid, s);
    }
    if (Cin == 64) {
      if (head) return launch_zf2<64, EPI_HEAD>(tx, tw, a, grid, s);
      return launch_zf2<64, EPI_STORE>(tx, tw, a, grid, s);
    }
    EXA_CHECK(!head, "conv_zfold: fused head needs Cin <= 64");
    return launch_zf2<128, EPI_STORE>(tx, tw, a, grid, s);
  }
  if (Cin == 32) {
    if (head) return launch_zf<32, EPI_HEAD>(tx, tw, a, grid, s);
    return launch_zf<32, EPI_STORE>(tx, tw, a, grid, s);
  }
  if (head) return launch_zf<64, EPI_HEAD>(tx, tw, a, grid, s);
  return launch_zf<64, EPI_STORE>(tx, tw, a, grid, s);
}

}  // namespace exa
