// K1s: the Cin = 1 stem conv (reference inc.double_conv.0 + BatchNorm + LeakyReLU,
// unet3d.py:143-145) on the tensor cores.
//
// A 1-channel 3x3x3 conv has K = 27: too thin for an implicit GEMM over channels, and 864 fp32
// FMAs per voxel on the SIMT pipes cost as much time (and more energy) as the 32->32 conv that
// follows.  Here the x axis is folded into the GEMM instead (Toeplitz form):
//   M = 128 rows = 8 (y) x 16 (z) voxel rows of the patch,
//   K = 16       = 8 consecutive input voxels x' = x0-1 .. x0+6 of one (z+kz-1, y+ky-1) row, each
//                  as a bf16 (hi, lo) pair,
//   N = 128      = 4 output voxels x0 .. x0+3  x  32 output channels,
//   B[kz,ky][n = (xo, c)][k = (x', part)] = w[c][kz][ky][x' - xo]  (band matrix, zero outside
//   0..2, the same weight for the hi and the lo part),
// so a tile of 512 voxels x 32 channels is 9 (kz,ky) MMAs of 64 cycles.  The A operand needs no
// im2col: the input patch is stored x-innermost as interleaved bf16 (hi, lo) pairs, one TMA box
// brings the (8 x', 10 y, 18 z) halo tile (32-byte rows), and the nine (kz,ky) shifts are nine
// views of it (descriptor start + (kz*10+ky) rows, 8-row groups 10 rows apart), exactly as in
// conv_zfold.cuh.
// TMA moves rows, and 32-byte rows cost as much to issue as 128-byte ones (measured: ~10 cycles
// per row, the first version of this kernel was bound by it).  One box therefore carries 32
// voxels per row (128 B, SWIZZLE_128B) and serves SIX consecutive x sub-tiles: sub-tile j reads
// the 32 bytes at column 16*j of every row (descriptor start + 16*j bytes).
//
// Precision: the normalised input (fp32 in [0,1]) is split as x = hi + lo (exact to 2^-17
// relative); both parts are multiplied by the bf16-rounded folded weights and accumulated in
// fp32.  The input is therefore NOT rounded to bf16; the weights are rounded once like those of
// every other layer.
#pragma once

#include "common.cuh"
#include "conv_zfold.cuh"  // descriptor helpers, st_global_256

namespace exa {

constexpr int STEM_SUBS = 6;  // x sub-tiles (4 outputs each) per TMA box: 24 outputs from 32 voxels

struct StemTcArgs {
  int B, P[3];               // patches, patch dims (z, y, x); multiples of 16 (z), 8 (y), 4 (x)
  int ntz, nty, ntx;         // super-tiles: 16 (z) x 8 (y) x 24 (x) output voxels (last x one ragged)
  int tiles_total;
  const float* bias;         // [32] folded
  __nv_bfloat16* out;        // NDHWC: 32 channels at offset coff of voxels cstride elements apart
  int cstride, coff;         // multiples of 16 (32-byte stores)
};

struct StemTcSmem {
  static constexpr int A_ROWS = 180;                 // (10 y) x (18 z) rows of 32 x' x (hi, lo) (128 B)
  static constexpr int A_TX_BYTES = A_ROWS * 128;
  static constexpr int A_STAGE = 23552;              // 23040 B rounded to 1024
  static constexpr int W_TAP = 128 * 32;             // band matrix of one (kz, ky)
  static constexpr int W_BYTES = 9 * W_TAP;
  static constexpr int STAGES = 3;
  static constexpr int ACCS = 4;                     // accumulator buffers of 128 TMEM columns
  static constexpr int BAR_BYTES = 256;
  static constexpr int TOTAL = W_BYTES + STAGES * A_STAGE + BAR_BYTES + 1024;
};

__global__ void __launch_bounds__(ZF_THREADS, 1)
stem_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
               const StemTcArgs p) {
  using S = StemTcSmem;
  constexpr int STAGES = S::STAGES;
  constexpr int ACCS = S::ACCS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + S::W_BYTES;
  uint64_t* bars = (uint64_t*)(smem_a + STAGES * S::A_STAGE);
  uint64_t* full_bar = bars;                       // [STAGES] TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;             // [STAGES] MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;         // [ACCS] MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * STAGES + ACCS; // [ACCS] epilogue -> MMA
  uint64_t* w_bar = bars + 2 * STAGES + 2 * ACCS;
  uint32_t* tmem_slot = (uint32_t*)(w_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < ACCS; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), 4);  // the four warps of the set that drains the tile
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
        // box (32 x' as 64 bf16, 10 y, 18 z, 1 b) at voxel (x0-1, y0-1, z0-1): rows are stored
        // shifted by one voxel, so the innermost coordinate is 2*x0 elements (16 B aligned);
        // out-of-patch y/z (and x beyond the padded row) are zero-filled by TMA
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_a + stage * S::A_STAGE)),
            "l"(reinterpret_cast<uint64_t>(&tmap_x)), "r"(fb), "r"(tx * (8 * STEM_SUBS)), "r"(ty * 8 - 1),
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
      constexpr uint32_t IDESC = umma_idesc_bf16(128, 128);
      const uint64_t w_desc = zf_join(zf_desc_lo(smem_u32(smem_w)), zf_desc_hi<32>(8));
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
        const int tx = tile % p.ntx;
        const int nsub = min(STEM_SUBS, (p.P[2] - tx * (4 * STEM_SUBS)) >> 2);
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        tc_fence_after();
        // halo view: rows are (z, y) with y innermost; 8-row groups (8 y of one z) 10 rows apart
        const uint64_t a_desc =
            zf_join(zf_desc_lo(smem_u32(smem_a + stage * S::A_STAGE)), zf_desc_hi<128>(10));
        for (int j = 0; j < nsub; ++j) {
          mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d0 = tmem_base + (uint32_t)(acc * 128);
#pragma unroll
          for (int t = 0; t < 9; ++t) {  // (kz, ky); sub-tile j starts 16*j bytes into every row
            const uint64_t ad = a_desc + (uint64_t)((((t / 3) * 10 + (t % 3)) * 128) >> 4) + (uint64_t)j;
            const uint64_t bd = w_desc + (uint64_t)((t * S::W_TAP) >> 4);
            umma_bf16(d0, ad, bd, IDESC, t == 0 ? 0u : 1u);
          }
          umma_commit(smem_u32(&tfull_bar[acc]));
          if (++acc == ACCS) {
            acc = 0;
            acc_phase ^= 1u;
          }
        }
        umma_commit(smem_u32(&empty_bar[stage]));
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue: two sets of 4 warps take alternate tiles =====================
    // The epilogue is latency-bound (barrier wake-up -> tcgen05.ld -> convert -> store), so the two
    // warp sets work on DIFFERENT tiles: warp (q, hs) handles rows 32q.. of every tile with
    // sequence parity hs and writes all four x outputs of its voxel row (256 contiguous bytes).
    const int q = warp & 3;
    const int hs = (warp - 4) >> 2;
    const int row = q * 32 + lane;           // row = zz * 8 + yy
    const int yy = row & 7, zz = row >> 3;
    float bias[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bias[j] = __ldg(p.bias + j);
    int acc = 0;
    uint32_t acc_phase = 0;
    int seq = 0;
    for (int tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
      const int b = tile / tiles_per_b;
      int r = tile - b * tiles_per_b;
      const int tz = r / (p.nty * p.ntx);
      r -= tz * (p.nty * p.ntx);
      const int ty = r / p.ntx, tx = r - ty * p.ntx;
      const int nsub = min(STEM_SUBS, (p.P[2] - tx * (4 * STEM_SUBS)) >> 2);
      const int z = tz * 16 + zz, y = ty * 8 + yy;
      __nv_bfloat16* row_dst =
          p.out + p.coff +
          ((((size_t)b * p.P[0] + z) * p.P[1] + y) * p.P[2] + tx * (4 * STEM_SUBS)) * p.cstride;
      for (int j = 0; j < nsub; ++j, ++seq) {
        if ((seq & 1) == hs) {
          __nv_bfloat16* dst = row_dst + (size_t)j * 4 * p.cstride;
          mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
          tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * 128);
#pragma unroll
          for (int half = 0; half < 2; ++half) {  // x outputs (0, 1) then (2, 3): 64 registers in flight
            uint32_t v0[32], v1[32];
            tmem_ld_32x32(taddr + (uint32_t)(half * 64), v0);
            tmem_ld_32x32(taddr + (uint32_t)(half * 64 + 32), v1);
            tmem_ld_wait();
            if (half == 1) {  // the accumulator is in registers: hand it back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
            }
#pragma unroll
            for (int xo = 0; xo < 2; ++xo) {
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                uint32_t pk[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                  const uint32_t ua = xo == 0 ? v0[16 * g + 2 * jj] : v1[16 * g + 2 * jj];
                  const uint32_t uc = xo == 0 ? v0[16 * g + 2 * jj + 1] : v1[16 * g + 2 * jj + 1];
                  const float a = leaky_relu(__uint_as_float(ua) + bias[16 * g + 2 * jj]);
                  const float c = leaky_relu(__uint_as_float(uc) + bias[16 * g + 2 * jj + 1]);
                  pk[jj] = pack_bf16x2(a, c);
                }
                st_global_256(dst + (size_t)(half * 2 + xo) * p.cstride + g * 16, pk);
              }
            }
          }
        }
        if (++acc == ACCS) {
          acc = 0;
          acc_phase ^= 1u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace exa
