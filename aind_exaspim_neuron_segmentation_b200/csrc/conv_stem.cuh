// K1s: the Cin = 1 stem conv (reference inc.double_conv.0 + BatchNorm + LeakyReLU,
// unet3d.py:143-145) on the tensor cores.
//
// A 1-channel 3x3x3 conv has K = 27: too thin for an implicit GEMM over channels, and 864 fp32
// FMAs per voxel on the SIMT pipes cost as much time (and more energy) as the 32->32 conv that
// follows.  Here the x axis is folded into the GEMM instead (Toeplitz form):
//   M = 128 rows  = 8 (y) x 16 (z) voxels rows of the patch,
//   K = 16        = 16 consecutive input voxels x' = x0-1 .. x0+14 of one (z+kz-1, y+ky-1) row,
//   N = 256       = 8 output voxels x0 .. x0+7  x  32 output channels,
//   B[kz,ky][n = (xo, c)][k = x'] = w[c][kz][ky][x' - xo]  (band matrix, zero outside 0..2),
// so a tile of 1024 voxels x 32 channels is 9 (kz,ky) MMAs.  The A operand needs no im2col: the
// input patch is stored x-innermost as bf16, one TMA box brings the (16 x', 10 y, 18 z) halo
// tile, and the nine (kz,ky) shifts are nine views of it (descriptor start + (kz*10+ky) rows,
// 8-row groups 10 rows apart), exactly as in conv_zfold.cuh.
//
// Precision: the normalised input (fp32 in [0,1]) is split into bf16 hi + lo parts
// (x = hi + lo up to 2^-17 relative), both multiplied by the bf16-rounded folded weights and
// accumulated in fp32: 18 MMAs per tile.  The input is therefore NOT rounded to bf16; the
// weights are rounded once like those of every other layer.
#pragma once

#include "common.cuh"
#include "conv_zfold.cuh"  // descriptor helpers, st_global_256

namespace exa {

struct StemTcArgs {
  int B, P[3];               // patches, patch dims (z, y, x); multiples of 16 (z), 8 (y, x)
  int ntz, nty, ntx;         // tiles: 16 (z) x 8 (y) x 8 (x) output voxels
  int tiles_total;
  const float* bias;         // [32] folded
  __nv_bfloat16* out;        // NDHWC, C = 32 dense
};

struct StemTcSmem {
  static constexpr int A_ROWS = 180;                 // (10 y) x (18 z) rows of 16 x' (32 B)
  static constexpr int A_PART = 6144;                // 5760 B rounded to 1024
  static constexpr int A_STAGE = 2 * A_PART;         // hi + lo
  static constexpr int A_TX_BYTES = 2 * A_ROWS * 32;
  static constexpr int W_TAP = 256 * 32;             // band matrix of one (kz, ky)
  static constexpr int W_BYTES = 9 * W_TAP;
  static constexpr int STAGES = 4;
  static constexpr int BAR_BYTES = 256;
  static constexpr int TOTAL = W_BYTES + STAGES * A_STAGE + BAR_BYTES + 1024;
};

__global__ void __launch_bounds__(ZF_THREADS, 1)
stem_tc_kernel(const __grid_constant__ CUtensorMap tmap_hi, const __grid_constant__ CUtensorMap tmap_lo,
               const __grid_constant__ CUtensorMap tmap_w, const StemTcArgs p) {
  using S = StemTcSmem;
  constexpr int STAGES = S::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + S::W_BYTES;
  uint64_t* bars = (uint64_t*)(smem_a + STAGES * S::A_STAGE);
  uint64_t* full_bar = bars;                    // [STAGES] TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;          // [STAGES] MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;      // [2] MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2] epilogue -> MMA
  uint64_t* w_bar = bars + 2 * STAGES + 4;
  uint32_t* tmem_slot = (uint32_t*)(w_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_hi);
    tma_prefetch_desc(&tmap_lo);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), 8);
    }
    mbar_init(smem_u32(w_bar), 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_b = p.ntz * p.nty * p.ntx;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_expect_tx(smem_u32(w_bar), (uint32_t)S::W_BYTES);
      for (int t = 0; t < 9; ++t)
        tma_load_3d(smem_u32(smem_w + t * S::W_TAP), &tmap_w, smem_u32(w_bar), 0, 0, t);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
        const int b = tile / tiles_per_b;
        int r = tile - b * tiles_per_b;
        const int tz = r / (p.nty * p.ntx);
        r -= tz * (p.nty * p.ntx);
        const int ty = r / p.ntx, tx = r - ty * p.ntx;
        mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
        const uint32_t fb = smem_u32(&full_bar[stage]);
        mbar_expect_tx(fb, (uint32_t)S::A_TX_BYTES);
        const uint32_t dst = smem_u32(smem_a + stage * S::A_STAGE);
        // box (16 x', 10 y, 18 z, 1 b) at (x0-1, y0-1, z0-1); rows are stored shifted by one element,
        // so the innermost coordinate is x0 (16 B aligned); out-of-patch y/z are zero-filled
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
            "l"(reinterpret_cast<uint64_t>(&tmap_hi)), "r"(fb), "r"(tx * 8), "r"(ty * 8 - 1),
            "r"(tz * 16 - 1), "r"(b)
            : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst + (uint32_t)S::A_PART),
            "l"(reinterpret_cast<uint64_t>(&tmap_lo)), "r"(fb), "r"(tx * 8), "r"(ty * 8 - 1),
            "r"(tz * 16 - 1), "r"(b)
            : "memory");
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      mbar_wait(smem_u32(w_bar), 0);
      tc_fence_after();
      constexpr uint32_t IDESC = umma_idesc_bf16(128, 256);
      const uint64_t w_desc = zf_join(zf_desc_lo(smem_u32(smem_w)), zf_desc_hi<32>(8));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      bool ftok = false, ttok = false;  // next tile's barriers already seen complete (probed early:
                                        // a probe consumed at once stalls the tensor pipe)
      for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
        if (!ttok) mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1u);
        if (!ftok) mbar_wait(smem_u32(&full_bar[stage]), phase);
        {
          const int ns = stage + 1 == STAGES ? 0 : stage + 1;
          const uint32_t np = stage + 1 == STAGES ? phase ^ 1u : phase;
          const int na = acc ^ 1;
          const uint32_t nap = na == 0 ? acc_phase ^ 1u : acc_phase;
          ftok = mbar_test_wait(smem_u32(&full_bar[ns]), np);
          ttok = mbar_test_wait(smem_u32(&tempty_bar[na]), nap ^ 1u);
        }
        tc_fence_after();
        // halo view: rows are (z, y) with y innermost; 8-row groups (8 y of one z) 10 rows apart
        const uint64_t a_desc =
            zf_join(zf_desc_lo(smem_u32(smem_a + stage * S::A_STAGE)), zf_desc_hi<32>(10));
        const uint32_t d0 = tmem_base + (uint32_t)(acc * 256);
#pragma unroll
        for (int part = 0; part < 2; ++part) {   // hi, lo
#pragma unroll
          for (int t = 0; t < 9; ++t) {          // (kz, ky)
            const uint64_t ad = a_desc + (uint64_t)((part * S::A_PART + ((t / 3) * 10 + (t % 3)) * 32) >> 4);
            const uint64_t bd = w_desc + (uint64_t)((t * S::W_TAP) >> 4);
            umma_bf16(d0, ad, bd, IDESC, (part == 0 && t == 0) ? 0u : 1u);
          }
        }
        umma_commit(smem_u32(&empty_bar[stage]));
        umma_commit(smem_u32(&tfull_bar[acc]));
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue: 8 warps; warp pair (q, hs) = rows 32q.., x outputs 4hs.. =====
    const int q = warp & 3;
    const int hs = (warp - 4) >> 2;
    const int row = q * 32 + lane;           // row = zz * 8 + yy
    const int yy = row & 7, zz = row >> 3;
    float bias[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bias[j] = __ldg(p.bias + j);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
      const int b = tile / tiles_per_b;
      int r = tile - b * tiles_per_b;
      const int tz = r / (p.nty * p.ntx);
      r -= tz * (p.nty * p.ntx);
      const int ty = r / p.ntx, tx = r - ty * p.ntx;
      const int z = tz * 16 + zz, y = ty * 8 + yy, x = tx * 8 + hs * 4;
      __nv_bfloat16* dst = p.out + ((((size_t)b * p.P[0] + z) * p.P[1] + y) * p.P[2] + x) * 32;
      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 256 + hs * 128);
#pragma unroll
      for (int xo = 0; xo < 4; ++xo) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + (uint32_t)(xo * 32), v);
        tmem_ld_wait();
        if (xo == 3) {  // the accumulator is in registers: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float a = leaky_relu(__uint_as_float(v[16 * g + 2 * j]) + bias[16 * g + 2 * j]);
            const float c = leaky_relu(__uint_as_float(v[16 * g + 2 * j + 1]) + bias[16 * g + 2 * j + 1]);
            pk[j] = pack_bf16x2(a, c);
          }
          st_global_256(dst + xo * 32 + g * 16, pk);
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace exa
