// T4tc: conv weight gradient on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
//   dW[co][ci][kz,ky,kx] = sum_v dz[v][co] * x[v + (kz-1, ky-1, kx-1)][ci]
//
// is a GEMM whose reduction dimension is the VOXEL index, while both operands are stored
// channels-innermost (NDHWC): both are "MN-major" for the tensor core.  tcgen05.mma reads
// MN-major bf16 operands directly (instruction-descriptor bits 15/16), and a TMA box
// [voxels][32 channels] written with SWIZZLE_64B is exactly the canonical MN-major layout
// (32 channels = one 64-byte row per K index, 8-row groups 512 B apart).  The descriptor's
// stride between 32-element MN blocks is free, which folds two of the three tap axes into
// ONE instruction:
//   * A (M = 128): four blocks 64 B apart = the x tile shifted by 0..3 voxels -> kx = 0, 1, 2
//     (the fourth block is padding: M must be 64 or 128);
//   * B (N = 96): three blocks one tile row apart = the dz rows y-1, y, y+1 -> ky = 2, 1, 0;
//   * kz: three accumulators of 96 TMEM columns, one per x plane z-1, z, z+1.
// One instruction (K = 16 voxels of a row) therefore carries 9 of the 27 taps of a 32 x 32
// (Cin x Cout) block; the accumulators stay in TMEM for the CTA's whole share of the volume
// and are written out once (split-K partial sums, reduced by launch_wgrad_reduce).
//
// A CTA marches columns (batch, 8 rows, 16 voxels of x) along z: per step one new x plane
// [8][20][32] and one dz plane [10][16][32] arrive by TMA (20 KB for 24 instructions), four
// steps ahead of their use.
// Warp 0: TMA producer, warp 1: MMA issuer, all four warps: final TMEM read-out.
#include <stdlib.h>

#include "tmap.h"
#include "train_kernels.h"

namespace exa {

namespace {

constexpr int WT_RY = 8;    // x rows per tile (dz rows: WT_RY + 2)
constexpr int WT_XW = 20;   // staged voxels per x row: 16 + one halo voxel each side + 2 pad
constexpr int WT_XS = 7;    // x plane slots: three live + four in flight (TMA latency is 2-3 steps)
constexpr int WT_DS = 5;    // dz plane slots: one live + four in flight
constexpr int WT_X_BYTES = WT_RY * WT_XW * 64;      // 10240
constexpr int WT_D_BYTES = (WT_RY + 2) * 16 * 64;   // 10240
constexpr int WT_BAR_OFF = WT_XS * WT_X_BYTES + WT_DS * WT_D_BYTES;
constexpr int WT_SMEM = 1024 + WT_BAR_OFF + 256;
constexpr uint32_t WT_TMEM_COLS = 512;              // 3 x 96 used

struct WtArgs {
  float* partial;  // [splits][cout][cin][27]
  int D, H, W, cin, cout;
  int nty, ntx, cols_total;
};

// MN-major operand, SWIZZLE_64B: 32 MN elements (64 B) per K row, K rows 64 B apart, 8-row K
// groups 512 B apart (SBO), 32-element MN blocks `lbo` bytes apart (LBO).
__device__ __forceinline__ uint64_t desc_mn64(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100)
  d |= (uint64_t)4 << 61;   // SWIZZLE_64B
  return d;
}

}  // namespace

__global__ void __launch_bounds__(128, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_d,
                const WtArgs a) {
  extern __shared__ uint8_t wt_smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)wt_smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* xs = smem;
  uint8_t* ds = smem + WT_XS * WT_X_BYTES;
  uint64_t* bars = (uint64_t*)(smem + WT_BAR_OFF);
  uint64_t* full_x = bars;                      // [WT_XS]  TMA -> MMA
  uint64_t* empty_x = bars + WT_XS;             // [WT_XS]  MMA -> TMA
  uint64_t* full_d = bars + 2 * WT_XS;          // [WT_DS]
  uint64_t* empty_d = bars + 2 * WT_XS + WT_DS;  // [WT_DS]
  uint64_t* done_bar = bars + 2 * WT_XS + 2 * WT_DS;
  uint32_t* tmem_slot = (uint32_t*)(done_bar + 1);
  uint32_t* issued = (uint32_t*)(done_bar + 2);  // [3] accumulator kz has received an MMA
  static_assert((2 * WT_XS + 2 * WT_DS + 4) * 8 <= 256, "barrier block");

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int cib = blockIdx.y, cob = blockIdx.z;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_d);
    for (int s = 0; s < WT_XS; ++s) {
      mbar_init(smem_u32(&full_x[s]), 1);
      mbar_init(smem_u32(&empty_x[s]), 1);
    }
    for (int s = 0; s < WT_DS; ++s) {
      mbar_init(smem_u32(&full_d[s]), 1);
      mbar_init(smem_u32(&empty_d[s]), 1);
    }
    mbar_init(smem_u32(done_bar), 1);
    issued[0] = issued[1] = issued[2] = 0u;
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), WT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int D = a.D;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t xg = 0, dg = 0;  // planes loaded so far (slot = count % slots)
      for (int col = blockIdx.x; col < a.cols_total; col += gridDim.x) {
        const int xt = col % a.ntx;
        const int yt = (col / a.ntx) % a.nty;
        const int b = col / (a.ntx * a.nty);
        const int x0 = xt * 16, y0 = yt * WT_RY;
        for (int z = 0; z < D; ++z) {
          // step z consumes x planes z-1, z, z+1 and dz plane z
          for (int zp = (z == 0 ? 0 : z + 1); zp <= z + 1 && zp < D; ++zp) {
            const uint32_t slot = xg % WT_XS;
            mbar_wait(smem_u32(&empty_x[slot]), ((xg / WT_XS) & 1u) ^ 1u);
            const uint32_t fb = smem_u32(&full_x[slot]);
            mbar_expect_tx(fb, (uint32_t)WT_X_BYTES);
            tma_load_5d(smem_u32(xs + slot * WT_X_BYTES), &tmap_x, fb, cib * 32, x0 - 1, y0, zp, b);
            ++xg;
          }
          const uint32_t slot = dg % WT_DS;
          mbar_wait(smem_u32(&empty_d[slot]), ((dg / WT_DS) & 1u) ^ 1u);
          const uint32_t fb = smem_u32(&full_d[slot]);
          mbar_expect_tx(fb, (uint32_t)WT_D_BYTES);
          tma_load_5d(smem_u32(ds + slot * WT_D_BYTES), &tmap_d, fb, cob * 32, x0, y0 - 1, z, b);
          ++dg;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 96) | (1u << 15) | (1u << 16);  // A, B MN-major
      uint32_t xbase = 0, dbase = 0;  // global index of this column's plane 0
      uint32_t first[3] = {1u, 1u, 1u};
      bool peek_x = false, peek_d = false;
      for (int col = blockIdx.x; col < a.cols_total; col += gridDim.x) {
        const int yt = (col / a.ntx) % a.nty;
        const int rows = min(WT_RY, a.H - yt * WT_RY);  // x rows inside the volume
        for (int z = 0; z < D; ++z) {
          // A barrier probe whose result is consumed at once stalls the issue stream for ~150
          // cycles (tools/umma_probe.cu); the planes of step z were probed during step z-1
          // (they arrive four steps ahead), so the blocking wait only runs when the probe failed.
          if (z == 0) {
            const uint32_t g = xbase;
            mbar_wait(smem_u32(&full_x[g % WT_XS]), (g / WT_XS) & 1u);
            peek_x = peek_d = false;
          }
          if (z + 1 < D && !peek_x) {
            const uint32_t g = xbase + (uint32_t)z + 1u;
            mbar_wait(smem_u32(&full_x[g % WT_XS]), (g / WT_XS) & 1u);
          }
          const uint32_t gd = dbase + (uint32_t)z;
          if (!peek_d) mbar_wait(smem_u32(&full_d[gd % WT_DS]), (gd / WT_DS) & 1u);
          peek_x = peek_d = false;
          if (z + 2 < D) {
            const uint32_t g = xbase + (uint32_t)z + 2u;
            peek_x = mbar_test_wait(smem_u32(&full_x[g % WT_XS]), (g / WT_XS) & 1u);
          }
          if (z + 1 < D) {
            const uint32_t g = gd + 1u;
            peek_d = mbar_test_wait(smem_u32(&full_d[g % WT_DS]), (g / WT_DS) & 1u);
          }
          tc_fence_after();
          const uint32_t da = smem_u32(ds + (gd % WT_DS) * WT_D_BYTES);
#pragma unroll
          for (int kz = 0; kz < 3; ++kz) {
            const int zp = z + kz - 1;
            if (zp < 0 || zp >= D) continue;  // zero plane: nothing to add
            const uint32_t g = xbase + (uint32_t)zp;
            const uint32_t xa = smem_u32(xs + (g % WT_XS) * WT_X_BYTES);
            for (int r = 0; r < rows; ++r) {
              // A: x row r, blocks = voxel shifts 0..3; B: dz rows r, r+1, r+2 of the tile
              // (= y-1, y, y+1), blocks one row (16 voxels) apart
              umma_bf16(tmem_base + (uint32_t)(kz * 96), desc_mn64(xa + (uint32_t)(r * WT_XW * 64), 64u),
                        desc_mn64(da + (uint32_t)(r * 16 * 64), 16u * 64u), idesc, first[kz] ^ 1u);
              first[kz] = 0u;
            }
          }
          // plane z-1 has had its last use (as kz = 0); the last step also retires plane D-1
          if (z >= 1) umma_commit(smem_u32(&empty_x[(xbase + (uint32_t)z - 1u) % WT_XS]));
          if (z == D - 1) umma_commit(smem_u32(&empty_x[(xbase + (uint32_t)z) % WT_XS]));
          umma_commit(smem_u32(&empty_d[gd % WT_DS]));
        }
        xbase += (uint32_t)D;
        dbase += (uint32_t)D;
      }
      issued[0] = first[0] ^ 1u;
      issued[1] = first[1] ^ 1u;
      issued[2] = first[2] ^ 1u;
      umma_commit(smem_u32(done_bar));
    }
    __syncwarp();
  }
  __syncthreads();  // publishes issued[]
  mbar_wait(smem_u32(done_bar), 0u);
  tc_fence_after();

  // ===================== read-out: lane = (kx, ci), column = (kz, 2 - ky, co) =====================
  if (warp < 3) {
    const int kx = warp;
    const int ci = cib * 32 + lane;
    float* dst = a.partial + (size_t)blockIdx.x * a.cout * a.cin * 27;
    for (int kz = 0; kz < 3; ++kz) {
      const bool have = issued[kz] != 0u;
      for (int bb = 0; bb < 3; ++bb) {
        uint32_t r[32];
        if (have) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kz * 96 + bb * 32), r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        const int tap = kz * 9 + (2 - bb) * 3 + kx;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          dst[((size_t)(cob * 32 + j) * a.cin + ci) * 27 + tap] = __uint_as_float(r[j]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, WT_TMEM_COLS);
}

bool wgrad_tc_enabled() {
  static const bool on = []() {
    const char* e = getenv("EXA_WGRAD");  // "mma": the warp-level mma.sync kernel (A/B, fallback)
    return !(e && e[0] == 'm');
  }();
  return on;
}

int wgrad_tc_columns(const Act& x) { return x.B * ceil_div(x.H, WT_RY) * ceil_div(x.W, 16); }

// splits: CTAs per (Cin, Cout) block, one wave of at most num_sms CTAs, columns spread evenly
int wgrad_tc_splits(const Act& x, int cout, int num_sms) {
  const int blocks = (x.C / 32) * (cout / 32);
  const int cols = wgrad_tc_columns(x);
  int smax = num_sms / blocks;
  if (smax < 1) smax = 1;
  if (smax > cols) smax = cols;
  const int per = ceil_div(cols, smax);
  return ceil_div(cols, per);
}

Status launch_wgrad_tc(const Act& x, const Act& dz, float* partial, int num_sms, cudaStream_t s) {
  EXA_CHECK(!x.fp32 && !dz.fp32 && x.C % 32 == 0 && dz.C % 32 == 0 && x.cstride % 8 == 0 &&
                x.coff % 8 == 0 && dz.cstride == dz.C && dz.coff == 0,
            "wgrad_tc: bf16 operands with channel counts that are multiples of 32");
  WtArgs a{};
  a.partial = partial;
  a.D = x.D; a.H = x.H; a.W = x.W; a.cin = x.C; a.cout = dz.C;
  a.nty = ceil_div(x.H, WT_RY);
  a.ntx = ceil_div(x.W, 16);
  a.cols_total = wgrad_tc_columns(x);
  CUtensorMap tx, td;
  {
    const uint64_t cs = (uint64_t)x.cstride * 2;
    uint64_t dims[5] = {(uint64_t)x.C, (uint64_t)x.W, (uint64_t)x.H, (uint64_t)x.D, (uint64_t)x.B};
    uint64_t strides[4] = {cs, cs * x.W, cs * x.W * x.H, cs * x.W * x.H * x.D};
    uint32_t box[5] = {32, (uint32_t)WT_XW, (uint32_t)WT_RY, 1, 1};
    void* base = (void*)((__nv_bfloat16*)x.ptr + x.coff);
    EXA_TRY(make_tmap_bf16(&tx, base, 5, dims, strides, box, 64));
  }
  {
    const uint64_t cs = (uint64_t)dz.C * 2;
    uint64_t dims[5] = {(uint64_t)dz.C, (uint64_t)dz.W, (uint64_t)dz.H, (uint64_t)dz.D, (uint64_t)dz.B};
    uint64_t strides[4] = {cs, cs * dz.W, cs * dz.W * dz.H, cs * dz.W * dz.H * dz.D};
    uint32_t box[5] = {32, 16, (uint32_t)(WT_RY + 2), 1, 1};
    EXA_TRY(make_tmap_bf16(&td, dz.ptr, 5, dims, strides, box, 64));
  }
  static bool configured = false;
  if (!configured) {
    EXA_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  WT_SMEM));
    configured = true;
  }
  dim3 grid((unsigned)wgrad_tc_splits(x, dz.C, num_sms), (unsigned)(x.C / 32), (unsigned)(dz.C / 32));
  wgrad_tc_kernel<<<grid, 128, WT_SMEM, s>>>(tx, td, a);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

}  // namespace exa
