// K1z: 3x3x3 Conv3d (+ folded BatchNorm bias + LeakyReLU) for the wide, shallow layers
// (Cout tile of 32, Cin 32 or 64) where the plain implicit GEMM (conv_umma.cuh) is bound by
// shared-memory operand reads and by one TMA box per tap.  Replaces reference
// unet3d.py:143-148 per layer; with EPI_HEAD also unet3d.py:318 (OutConv), inference.py:158
// (sigmoid) and inference.py:161-162 (trim).
//
// B200-first design (measured motivation in DESIGN.md):
//  * A CTA owns one 8(x) x 16(y) voxel column of one patch and marches along z.  Per input
//    plane ONE TMA box load brings the (8+2) x (16+2) halo tile of all Cin channels into shared
//    memory (out-of-bounds coordinates are zero-filled = the conv's zero padding).  The nine
//    in-plane taps are nine views of that tile: the UMMA descriptor start address moves by
//    (ky*10 + kx) rows and the 8-row-group stride is 10 rows, so no data is re-fetched.
//  * The three z taps are folded into the MMA's N dimension: B = [W(kz=2); W(kz=1); W(kz=0)]
//    (96 rows), so one M=128,N=96 MMA adds input plane z's contribution to output planes
//    z-1, z, z+1 at once.  The accumulators of the 16 most recent output planes live in a
//    ring of 32-column slots in TMEM (all 512 columns); a plane is complete after input
//    plane z+1 and is drained by the epilogue warps while the MMAs continue.
//  * All 27 taps of the weights stay resident in shared memory for the CTA's lifetime.
#pragma once

#include "common.cuh"
#include "conv_umma.cuh"  // ConvEpilogue

namespace exa {

struct ZfArgs {
  int B, D, H, W;            // activation geometry (voxels)
  int ox, oy, oz;            // origin of the computed output region
  int ntx, nty, nzp;         // tiles along x (8 wide), y (16 tall), output planes
  int n_halves;              // Cout / 32 (each CTA computes one 32-channel slice)
  int tiles_total;           // B * nty * ntx
  const float* bias;         // [Cout] folded BN bias
  __nv_bfloat16* out;        // EPI_STORE: NDHWC bf16 output
  int out_cstride, out_coff;
  __nv_bfloat16* pool_out;   // optional fused MaxPool3d(2) output (unet3d.py:195), NDHWC at half res
  int pool_cstride, pool_coff;
  const float* head_w;       // EPI_HEAD: [head_c][32]
  const float* head_b;       // [head_c]
  float* head_out;           // [B][head_c][D-2t][H-2t][W-2t]
  int head_c, trim, apply_sigmoid;
  long long* dbg;            // optional [gridDim.x][8] cycle counters (EXA_ZF_DEBUG=1), else nullptr
};

template <int CIN>
struct ZfSmem {
  static constexpr int ROWB = CIN * 2;                          // bytes per voxel row
  static constexpr int A_ROWS = 180;                            // 10 x 18 halo tile
  static constexpr int A_TX_BYTES = A_ROWS * ROWB;
  static constexpr int A_STAGE = (A_TX_BYTES + 1023) / 1024 * 1024;
  static constexpr int W_TAP = 96 * ROWB;                       // [3 kz][32 cout] rows per tap
  static constexpr int W_BYTES = 9 * W_TAP;
  static constexpr int STAGES = CIN == 64 ? 5 : 8;
  static constexpr int BAR_BYTES = 512;
  static constexpr int TOTAL = W_BYTES + STAGES * A_STAGE + BAR_BYTES + 1024;
};

constexpr int ZF_RING = 16;  // TMEM slots of 32 fp32 columns

template <int CIN, int EPI>
__global__ void __launch_bounds__(256, 1)
conv3x3_zfold_kernel(const __grid_constant__ CUtensorMap tmap_x,
                     const __grid_constant__ CUtensorMap tmap_w, const ZfArgs p) {
  using S = ZfSmem<CIN>;
  constexpr int ROWB = S::ROWB;
  constexpr int STAGES = S::STAGES;
  constexpr int KSTEPS = CIN / 16;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + S::W_BYTES;
  uint64_t* bars = (uint64_t*)(smem_a + STAGES * S::A_STAGE);
  uint64_t* full_bar = bars;                          // [STAGES] TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;                // [STAGES] MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;            // [16] MMA -> epilogue (slot complete)
  uint64_t* tempty_bar = bars + 2 * STAGES + ZF_RING; // [16] epilogue -> MMA (slot drained)
  uint64_t* w_bar = bars + 2 * STAGES + 2 * ZF_RING;  // weights resident
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
    for (int s = 0; s < ZF_RING; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), 4);
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

  // work split: CTA parity selects the 32-channel output slice, so that the resident weights
  // never change; tiles are dealt round-robin over the CTAs of one parity class.
  const int half = blockIdx.x % p.n_halves;
  const int cta_in_class = blockIdx.x / p.n_halves;
  const int ctas_per_class = gridDim.x / p.n_halves;
  const int tiles_per_b = p.nty * p.ntx;
  const int zin0 = max(p.oz - 1, 0);
  const int zin1 = min(p.oz + p.nzp + 1, p.D);

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      // resident weights: 9 taps x [3 kz][32 cout][CIN]
      mbar_expect_tx(smem_u32(w_bar), (uint32_t)S::W_BYTES);
      for (int t = 0; t < 9; ++t) {
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_w + t * S::W_TAP)),
            "l"(reinterpret_cast<uint64_t>(&tmap_w)), "r"(smem_u32(w_bar)), "r"(0),
            "r"(half * 32), "r"(0), "r"(t)
            : "memory");
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cta_in_class; tile < p.tiles_total; tile += ctas_per_class) {
        const int b = tile / tiles_per_b;
        const int r = tile - b * tiles_per_b;
        const int ty = r / p.ntx, tx = r - ty * p.ntx;
        const int x0 = p.ox + tx * 8, y0 = p.oy + ty * 16;
        for (int zi = zin0; zi < zin1; ++zi) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, (uint32_t)S::A_TX_BYTES);
          tma_load_5d(smem_u32(smem_a + stage * S::A_STAGE), &tmap_x, fb, 0, x0 - 1, y0 - 1, zi, b);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      mbar_wait(smem_u32(w_bar), 0);
      tc_fence_after();
      const uint32_t w_base = smem_u32(smem_w);
      // Flattened loop over (tile, input plane).  The barrier probes for iteration it+1 are
      // issued BEFORE the MMAs of iteration it ("peek") and only consumed afterwards: an
      // mbarrier probe whose result is needed immediately costs 200-550 cycles, during which the
      // tensor pipe (which queues almost nothing) would drain.
      const int nin = zin1 - zin0;
      const int my_tiles = cta_in_class < p.tiles_total
                               ? (p.tiles_total - cta_in_class + ctas_per_class - 1) / ctas_per_class
                               : 0;
      const int total_it = my_tiles * nin;
      long long d_tempty = 0, d_full = 0, d_issue = 0, d_commit = 0, d_planes = 0;
      bool tok_full = false, tok_fresh = false;
      int zi = zin0;
      uint32_t gbase = 0;  // running count of output planes handled by this CTA
      for (int it = 0; it < total_it; ++it) {
        {
          const int stage = it % STAGES;
          const uint32_t phase = (uint32_t)(it / STAGES) & 1u;
          const long long c0 = p.dbg ? clock64() : 0;
          // Output planes fed by input plane zi: po = zi - 1 + kzr (kzr = 0..2 <-> B rows
          // [32*kzr, 32*kzr+32)).  Everything below is kept in scalar registers and the tap loop
          // is fully unrolled: a single thread issues every MMA, so its instruction count per
          // MMA bounds the tensor pipe.
          uint32_t sa_brow = 0, sa_n = 0, sa_col = 0;  // accumulate segment A (before a ring wrap)
          uint32_t sb_brow = 0, sb_n = 0, sb_col = 0;  // accumulate segment B (after a ring wrap)
          uint32_t f_col[3], f_acc[3], f_valid[3];
#pragma unroll
          for (int kzr = 0; kzr < 3; ++kzr) {
            const int po = zi - 1 + kzr;
            const bool valid = po >= p.oz && po < p.oz + p.nzp;
            const uint32_t g = gbase + (uint32_t)(po - p.oz);
            const uint32_t slot = g % ZF_RING;
            const bool fresh = zi == max(po - 1, zin0);
            f_valid[kzr] = valid ? 1u : 0u;
            f_col[kzr] = slot * 32u;
            f_acc[kzr] = fresh ? 0u : 1u;
            if (valid) {
              // first touch of this ring slot: the epilogue must have drained it.  The slot of
              // plane zi+1 was peeked during the previous iteration.
              if (fresh && !(kzr == 2 && tok_fresh)) {
                mbar_wait(smem_u32(&tempty_bar[slot]), ((g / ZF_RING) & 1u) ^ 1u);
              }
              if (sa_n == 0) {
                sa_brow = kzr * 32u; sa_n = 32u; sa_col = slot * 32u;
              } else if (sb_n == 0 && slot != 0) {
                sa_n += 32u;
              } else if (sb_n == 0) {
                sb_brow = kzr * 32u; sb_n = 32u; sb_col = slot * 32u;
              } else {
                sb_n += 32u;
              }
            }
          }
          const long long c1 = p.dbg ? clock64() : 0;
          if (!tok_full) mbar_wait(smem_u32(&full_bar[stage]), phase);
          const long long c2 = p.dbg ? clock64() : 0;
          // ---- peek the barriers of the next iteration ----
          int zi_next = zi + 1;
          uint32_t gbase_next = gbase;
          if (zi_next == zin1) {
            zi_next = zin0;
            gbase_next += (uint32_t)p.nzp;
          }
          tok_full = false;
          tok_fresh = false;
          if (it + 1 < total_it) {
            const int nstage = (it + 1) % STAGES;
            tok_full = mbar_test_wait(smem_u32(&full_bar[nstage]), (uint32_t)((it + 1) / STAGES) & 1u);
            const int pf = zi_next + 1;  // the plane first touched by the next iteration (kzr = 2)
            if (pf >= p.oz && pf < p.oz + p.nzp) {
              const uint32_t gn = gbase_next + (uint32_t)(pf - p.oz);
              tok_fresh = mbar_test_wait(smem_u32(&tempty_bar[gn % ZF_RING]),
                                         ((gn / ZF_RING) & 1u) ^ 1u);
            }
          }
          const uint32_t a_base = smem_u32(smem_a + stage * S::A_STAGE);
          // halo view: 8-row groups are 10 rows apart; tap (ky,kx) shifts the start by ky*10+kx rows
          uint64_t adesc0 = umma_smem_desc<ROWB>(a_base);
          adesc0 &= ~((uint64_t)0x3FFF << 32);
          adesc0 |= (uint64_t)((10 * ROWB) >> 4) << 32;
          const uint64_t bdesc0 = umma_smem_desc<ROWB>(w_base);
          // first k-step of tap 0: per-kz MMAs so that untouched slots are overwritten
#pragma unroll
          for (int kzr = 0; kzr < 3; ++kzr) {
            if (f_valid[kzr]) {
              umma_bf16(tmem_base + f_col[kzr], adesc0, bdesc0 + (uint64_t)((kzr * 32 * ROWB) >> 4),
                        umma_idesc_bf16(128, 32), f_acc[kzr]);
            }
          }
          const uint64_t bdA = bdesc0 + (uint64_t)((sa_brow * ROWB) >> 4);
          const uint64_t bdB = bdesc0 + (uint64_t)((sb_brow * ROWB) >> 4);
          const uint32_t idA = umma_idesc_bf16(128, (int)sa_n);
          const uint32_t idB = umma_idesc_bf16(128, (int)sb_n);
          const uint32_t dA = tmem_base + sa_col, dB = tmem_base + sb_col;
          if (sb_n == 0) {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                if (t == 0 && k == 0) continue;
                const uint64_t aoff = (uint64_t)((((t / 3) * 10 + (t % 3)) * ROWB + k * 32) >> 4);
                const uint64_t boff = (uint64_t)((t * S::W_TAP + k * 32) >> 4);
                umma_bf16(dA, adesc0 + aoff, bdA + boff, idA, 1u);
              }
            }
          } else {
#pragma unroll
            for (int t = 0; t < 9; ++t) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) {
                if (t == 0 && k == 0) continue;
                const uint64_t aoff = (uint64_t)((((t / 3) * 10 + (t % 3)) * ROWB + k * 32) >> 4);
                const uint64_t boff = (uint64_t)((t * S::W_TAP + k * 32) >> 4);
                umma_bf16(dA, adesc0 + aoff, bdA + boff, idA, 1u);
                umma_bf16(dB, adesc0 + aoff, bdB + boff, idB, 1u);
              }
            }
          }
          const long long c3 = p.dbg ? clock64() : 0;
          umma_commit(smem_u32(&empty_bar[stage]));
          // completed output planes: zi-1 always; zi too when it is the last input plane
          const int pc = zi - 1;
          if (pc >= p.oz && pc < p.oz + p.nzp)
            umma_commit(smem_u32(&tfull_bar[(gbase + (uint32_t)(pc - p.oz)) % ZF_RING]));
          if (zi == zin1 - 1 && zi >= p.oz && zi < p.oz + p.nzp)
            umma_commit(smem_u32(&tfull_bar[(gbase + (uint32_t)(zi - p.oz)) % ZF_RING]));
          if (p.dbg) {
            const long long c4 = clock64();
            d_tempty += c1 - c0; d_full += c2 - c1; d_issue += c3 - c2; d_commit += c4 - c3;
            ++d_planes;
          }
          zi = zi_next;
          gbase = gbase_next;
        }
      }
      if (p.dbg) {
        long long* d = p.dbg + (size_t)blockIdx.x * 8;
        d[0] = d_tempty; d[1] = d_full; d[2] = d_issue; d[3] = d_commit; d[4] = d_planes;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue: 4 warps, one TMEM lane quarter each =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int rx = row & 7, ry = row >> 3;
    const int n0 = half * 32;
    float bias[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) bias[j] = __ldg(p.bias + n0 + j);
    uint32_t gbase = 0;
    long long e_wait = 0, e_ld = 0, e_rest = 0;
    uint32_t prev[16];  // previous (even) plane, packed bf16x2, for the fused 2x2x2 max-pool
#pragma unroll
    for (int j = 0; j < 16; ++j) prev[j] = 0u;
    for (int tile = cta_in_class; tile < p.tiles_total; tile += ctas_per_class) {
      const int b = tile / tiles_per_b;
      const int r = tile - b * tiles_per_b;
      const int ty = r / p.ntx, tx = r - ty * p.ntx;
      const int x = p.ox + tx * 8 + rx, y = p.oy + ty * 16 + ry;
      const bool in_xy = x < p.W && y < p.H;
      for (int po = p.oz; po < p.oz + p.nzp; ++po) {
        const uint32_t g = gbase + (uint32_t)(po - p.oz);
        const uint32_t slot = g % ZF_RING;
        const long long e0 = p.dbg ? clock64() : 0;
        mbar_wait(smem_u32(&tfull_bar[slot]), (g / ZF_RING) & 1u);
        tc_fence_after();
        const long long e1 = p.dbg ? clock64() : 0;
        uint32_t acc[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + slot * 32u, acc);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[slot]));  // slot is in registers now
        const long long e2 = p.dbg ? clock64() : 0;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = leaky_relu(__uint_as_float(acc[j]) + bias[j]);
        if constexpr (EPI == EPI_STORE) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
          if (in_xy) {
            const size_t vox = (((size_t)b * p.D + po) * p.H + y) * p.W + x;
            uint4* dst = reinterpret_cast<uint4*>(p.out + vox * p.out_cstride + p.out_coff + n0);
#pragma unroll
            for (int gq = 0; gq < 4; ++gq)
              dst[gq] = make_uint4(pk[4 * gq], pk[4 * gq + 1], pk[4 * gq + 2], pk[4 * gq + 3]);
          }
          if (p.pool_out != nullptr) {
            // max over the z pair (this thread), the x pair (lane ^ 1) and the y pair (lane ^ 8);
            // max commutes with the bf16 rounding, so this equals pooling the stored tensor
            if (((po - p.oz) & 1) == 0) {
#pragma unroll
              for (int j = 0; j < 16; ++j) prev[j] = pk[j];
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                __nv_bfloat162 m = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&pk[j]),
                                           *reinterpret_cast<__nv_bfloat162*>(&prev[j]));
                uint32_t mu = *reinterpret_cast<uint32_t*>(&m);
                uint32_t o1 = __shfl_xor_sync(0xffffffffu, mu, 1);
                m = __hmax2(m, *reinterpret_cast<__nv_bfloat162*>(&o1));
                mu = *reinterpret_cast<uint32_t*>(&m);
                uint32_t o8 = __shfl_xor_sync(0xffffffffu, mu, 8);
                m = __hmax2(m, *reinterpret_cast<__nv_bfloat162*>(&o8));
                pk[j] = *reinterpret_cast<uint32_t*>(&m);
              }
              if (in_xy && (lane & 9) == 0) {
                const size_t pvox = (((size_t)b * (p.D >> 1) + (po >> 1)) * (p.H >> 1) + (y >> 1)) *
                                        (p.W >> 1) + (x >> 1);
                uint4* dst = reinterpret_cast<uint4*>(p.pool_out + pvox * p.pool_cstride +
                                                      p.pool_coff + n0);
#pragma unroll
                for (int gq = 0; gq < 4; ++gq)
                  dst[gq] = make_uint4(pk[4 * gq], pk[4 * gq + 1], pk[4 * gq + 2], pk[4 * gq + 3]);
              }
            }
          }
        } else {
          const int t = p.trim;
          const int Dz = p.D - 2 * t, Hy = p.H - 2 * t, Wx = p.W - 2 * t;
          const bool keep = in_xy && x >= t && x < p.W - t && y >= t && y < p.H - t && po >= t &&
                            po < p.D - t;
          if (keep) {
            for (int oc = 0; oc < p.head_c; ++oc) {
              float s = __ldg(p.head_b + oc);
#pragma unroll
              for (int j = 0; j < 32; ++j) s = fmaf(__ldg(p.head_w + oc * 32 + j), v[j], s);
              if (p.apply_sigmoid) s = 1.f / (1.f + expf(-s));
              const size_t o =
                  ((((size_t)b * p.head_c + oc) * Dz + (po - t)) * Hy + (y - t)) * Wx + (x - t);
              p.head_out[o] = s;
            }
          }
        }
        if (p.dbg) {
          const long long e3 = clock64();
          e_wait += e1 - e0; e_ld += e2 - e1; e_rest += e3 - e2;
        }
      }
      gbase += (uint32_t)p.nzp;
    }
    if (p.dbg && warp == 4 && lane == 0) {
      long long* d = p.dbg + (size_t)blockIdx.x * 8;
      d[5] = e_wait; d[6] = e_ld; d[7] = e_rest;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace exa
