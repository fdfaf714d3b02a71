// K1z2: the z-folded halo-tile conv of conv_zfold.cuh on CTA PAIRS (tcgen05 cta_group::2).
// Replaces reference unet3d.py:143-148 per layer (Conv3d k=3 p=1 + BatchNorm3d(eval) +
// LeakyReLU); with EPI_HEAD also unet3d.py:318, inference.py:158,161-162.
//
// Why pairs (measured): with one CTA per MMA the M=128,N=96 MMA is bound by
// shared-memory operand reads (A 32 + B 24 = 56 cycles against a 48-cycle tensor floor) and the
// 27 taps x 96 rows of resident weights limit Cin to 64.  A cta_group::2 MMA (M=256) lets the
// two SMs of a pair work on two different voxel tiles while each holds only HALF of the weight
// rows (N/2 = 48): B reads drop to 12 cycles (A 32 + B 12 < 48 -> tensor-bound), and the resident
// weights halve, which admits Cin = 128 (up3.conv.double_conv.0, the concat layer).
//
// Structure per CTA is that of conv_zfold.cuh (TMA producer warp, one MMA-issuing thread,
// 8 epilogue warps, per-input-plane accumulator groups in TMEM), with these differences:
//  * only the leader CTA (cluster rank 0) issues MMAs; its barriers collect the TMA bytes of both
//    CTAs (cta_group::2 TMA signals the leader's mbarrier) and the accumulator releases of both
//    epilogues (remote mbarrier.arrive); tcgen05.commit multicasts to both CTAs' barriers;
//  * Cin = 128 rows (256 B) exceed the 128 B swizzle span, so activations and weights are kept
//    as two 64-channel halves (two TMA boxes per plane, k-steps 0-3 / 4-7).
#pragma once

#include "common.cuh"
#include "conv_umma.cuh"   // ConvEpilogue
#include "conv_zfold.cuh"  // ZfArgs, descriptor helpers, tmem_ld_32x16, st_global_256

namespace exa {

template <int CIN>
struct Zf2Smem {
  static constexpr int KH = CIN == 128 ? 2 : 1;                 // 64-channel halves of K
  static constexpr int ROWB = CIN / KH * 2;                     // bytes per row of one half
  static constexpr int KPH = ROWB / 32;                         // k-steps (K=16) per half
  static constexpr int A_ROWS = 180;                            // 10 x 18 halo tile
  static constexpr int A_HALF_TX = A_ROWS * ROWB;
  static constexpr int A_HALF = (A_HALF_TX + 1023) / 1024 * 1024;
  static constexpr int A_STAGE = KH * A_HALF;
  static constexpr int A_TX_BYTES = KH * A_HALF_TX;             // per CTA
  static constexpr int W_HALF = 48 * ROWB;                      // this CTA's 48 of the 96 rows
  static constexpr int W_TAP = KH * W_HALF;
  static constexpr int W_BYTES = 9 * W_TAP;                     // per CTA
  static constexpr int STAGES = CIN == 128 ? 2 : (CIN == 64 ? 6 : 8);
  static constexpr int BAR_BYTES = 512;
  static constexpr int TOTAL = W_BYTES + STAGES * A_STAGE + BAR_BYTES + 1024;
};

constexpr uint32_t ZF2_PEER_MASK = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the
                                                 // even (leader) CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same offset in the leader CTA (local when we are the leader)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & ZF2_PEER_MASK) : "memory");
}
// TMA loads of a pair: data lands in the issuing CTA, the bytes are counted on the leader's barrier
__device__ __forceinline__ void tma2_load_5d(uint32_t dst, const void* tmap, uint32_t bar, int c0,
                                             int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar & ZF2_PEER_MASK), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0,
                                             int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar & ZF2_PEER_MASK), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of each CTA] * B[smem halves of both CTAs]; M = 256
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all MMAs issued so far retire) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit_both(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::
          "r"(bar),
      "h"((uint16_t)3)
      : "memory");
}

template <int CIN>
__device__ __forceinline__ void zf2_issue_plane(uint32_t d0, uint64_t a_desc, uint64_t w_desc) {
  using S = Zf2Smem<CIN>;
  constexpr int ROWB = S::ROWB;
  constexpr uint32_t I96 = umma_idesc_bf16(256, 96);
#pragma unroll
  for (int t = 0; t < 9; ++t) {
#pragma unroll
    for (int kh = 0; kh < S::KH; ++kh) {
#pragma unroll
      for (int k = 0; k < S::KPH; ++k) {
        const uint64_t ad = a_desc + (uint64_t)((kh * S::A_HALF + ((t / 3) * 10 + (t % 3)) * ROWB + k * 32) >> 4);
        const uint64_t bd = w_desc + (uint64_t)((t * S::W_TAP + kh * S::W_HALF + k * 32) >> 4);
        umma2_bf16(d0, ad, bd, I96, (t == 0 && kh == 0 && k == 0) ? 0u : 1u);
      }
    }
  }
}

template <int CIN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ZF_THREADS, 1)
conv3x3_zfold2_kernel(const __grid_constant__ CUtensorMap tmap_x,
                      const __grid_constant__ CUtensorMap tmap_w, const ZfArgs p) {
  using S = Zf2Smem<CIN>;
  constexpr int STAGES = S::STAGES;
  constexpr int ROWB = S::ROWB;
  constexpr uint32_t TEMPTY_COUNT = 16;  // 8 epilogue warps of each CTA release every group once

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + S::W_BYTES;
  uint64_t* bars = (uint64_t*)(smem_a + STAGES * S::A_STAGE);
  uint64_t* full_bar = bars;                             // [STAGES] TMA (both CTAs) -> MMA   (leader's is used)
  uint64_t* empty_bar = bars + STAGES;                   // [STAGES] MMA -> TMA              (each CTA's own)
  uint64_t* tfull_bar = bars + 2 * STAGES;               // [5] MMA -> epilogue              (each CTA's own)
  uint64_t* tempty_bar = bars + 2 * STAGES + ZF_GROUPS;  // [5] epilogues (both CTAs) -> MMA (leader's is used)
  uint64_t* w_bar = bars + 2 * STAGES + 2 * ZF_GROUPS;   // weights of both CTAs resident    (leader's is used)
  uint32_t* tmem_slot = (uint32_t*)(w_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < ZF_GROUPS; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), TEMPTY_COUNT);
    }
    mbar_init(smem_u32(w_bar), 1);
    mbar_fence_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything can signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work split: a PAIR owns one 32-channel output slice (class) and takes two consecutive tiles
  // per round, one per CTA; pairs of a class share the rounds round-robin.
  const int pair = blockIdx.x >> 1;
  const int half = pair % p.n_halves;
  const int pair_in_class = pair / p.n_halves;
  const int pairs_per_class = (gridDim.x >> 1) / p.n_halves;
  const int rounds_total = (p.tiles_total + 1) >> 1;
  const int my_rounds = pair_in_class < rounds_total
                            ? (rounds_total - pair_in_class + pairs_per_class - 1) / pairs_per_class
                            : 0;
  const int tiles_per_b = p.nty * p.ntx;
  const int zin0 = max(p.oz - 1, 0);
  const int zin1 = min(p.oz + p.nzp + 1, p.D);
  const int nin = zin1 - zin0;
  const int zend = p.oz + p.nzp;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (elect_one()) {
      // this CTA's 48 of the 96 [kz][cout] weight rows per tap, in chunks of 16 rows
      if (rank == 0) mbar_expect_tx(smem_u32(w_bar), 2u * (uint32_t)S::W_BYTES);
      for (int t = 0; t < 9; ++t) {
        for (int kh = 0; kh < S::KH; ++kh) {
          for (int j = 0; j < 3; ++j) {
            const int row = 48 * (int)rank + 16 * j;  // row in the 96-row [kz][32 cout] block
            tma2_load_4d(smem_u32(smem_w + t * S::W_TAP + kh * S::W_HALF + j * 16 * ROWB), &tmap_w,
                         smem_u32(w_bar), kh * 64, half * 32 + (row & 31), row >> 5, t);
          }
        }
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int rd = 0; rd < my_rounds; ++rd) {
        const int tile = 2 * (pair_in_class + rd * pairs_per_class) + (int)rank;
        // a pair's odd tile out: coordinates beyond the batch are zero-filled by TMA
        const int b = tile < p.tiles_total ? tile / tiles_per_b : p.B;
        const int r = tile < p.tiles_total ? tile - b * tiles_per_b : 0;
        const int ty = r / p.ntx, tx = r - ty * p.ntx;
        const int x0 = p.ox + tx * 8, y0 = p.oy + ty * 16;
        for (int zi = zin0; zi < zin1; ++zi) {
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          if (rank == 0) mbar_expect_tx(fb, 2u * (uint32_t)S::A_TX_BYTES);
          for (int kh = 0; kh < S::KH; ++kh)
            tma2_load_5d(smem_u32(smem_a + stage * S::A_STAGE + kh * S::A_HALF), &tmap_x, fb, kh * 64,
                         x0 - 1, y0 - 1, zi, b);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && elect_one()) {
      mbar_wait(smem_u32(w_bar), 0);
      tc_fence_after();
      const uint64_t w_desc_c = zf_join(zf_desc_lo(smem_u32(smem_w)), zf_desc_hi<ROWB>(8));
      const uint64_t a_desc0 = zf_join(zf_desc_lo(smem_u32(smem_a)), zf_desc_hi<ROWB>(10));
      const uint64_t a_desc_end = a_desc0 + (uint64_t)STAGES * (uint64_t)(S::A_STAGE >> 4);
      const uint32_t bar0 = smem_u32(bars);
      const uint32_t empty0 = bar0 + (uint32_t)STAGES * 8u;
      const uint32_t tfull0 = bar0 + 2u * (uint32_t)STAGES * 8u;
      const uint32_t tempty0 = tfull0 + (uint32_t)ZF_GROUPS * 8u;
      const uint32_t tempty_end = tempty0 + (uint32_t)ZF_GROUPS * 8u;
      uint64_t a_desc = a_desc0;
      uint32_t ea = empty0;
      uint32_t d0 = tmem_base;
      uint32_t tfa = tfull0;
      uint32_t fa = bar0, fph = 0;
      uint32_t ta = tempty0, tph = 1;
      bool ftok = false, ttok = false;

      long long dbg_c0 = 0, dbg_t0 = 0;
      if (p.dbg & 8) {
        dbg_c0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
      }
      const int total = my_rounds * nin;
      for (int it = 0; it < total; ++it) {
        if (!ftok) mbar_wait(fa, fph);
        if (!ttok) mbar_wait(ta, tph);
        fa += 8u;
        if (fa == empty0) {
          fa = bar0;
          fph ^= 1u;
        }
        ta += 8u;
        if (ta == tempty_end) {
          ta = tempty0;
          tph ^= 1u;
        }
        ftok = mbar_test_wait(fa, fph);
        ttok = mbar_test_wait(ta, tph);
        tc_fence_after();
        uint64_t w_desc = w_desc_c;
        asm volatile("" : "+l"(w_desc));
        zf2_issue_plane<CIN>(d0, a_desc, w_desc);
        umma2_commit_both(ea);   // shared-memory stage consumed (both CTAs' producers)
        umma2_commit_both(tfa);  // accumulator group complete (both CTAs' epilogues)
        ea += 8u;
        a_desc += (uint64_t)(S::A_STAGE >> 4);
        if (a_desc == a_desc_end) {
          a_desc = a_desc0;
          ea = empty0;
        }
        tfa += 8u;
        d0 += 96u;
        if (tfa == tempty0) {
          tfa = tfull0;
          d0 = tmem_base;
        }
      }
      if (p.dbg & 8) {
        long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        long long* d = p.dbg_out + (size_t)blockIdx.x * 4;
        d[0] = clock64() - dbg_c0;
        d[1] = t1 - dbg_t0;
        d[2] = (long long)total;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue: 8 warps per CTA, two per TMEM lane quarter =====================
    const int q = warp & 3;
    const int hs = (warp - 4) >> 2;
    const int row = q * 32 + lane;
    const int rx = row & 7, ry = row >> 3;
    const uint32_t tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar);
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t col_off = EPI == EPI_STORE ? (uint32_t)(hs * 16) : 0u;

    // wait until every partial of output plane po is complete
    auto wait_plane = [&](uint32_t gi0, int po) {
      const int zl = po + 1 < zin1 ? po + 1 : po;  // latest contributing input plane
      const uint32_t gi = gi0 + (uint32_t)(zl - zin0);
      mbar_wait(tfull0 + (gi % ZF_GROUPS) * 8u, (gi / ZF_GROUPS) & 1u);
      tc_fence_after();
    };
    // sum the (up to three) partials of 16 accumulator columns [c0, c0+16) of output plane po:
    // all loads are issued before the single wait
    auto gather16 = [&](uint32_t gi0, int po, uint32_t c0, float (&v)[16]) {
      uint32_t a[3][16];
#pragma unroll
      for (int dz = -1; dz <= 1; ++dz) {
        const int z = po + dz;
        if (z >= zin0 && z < zin1) {
          const uint32_t gi = gi0 + (uint32_t)(z - zin0);
          tmem_ld_32x16(tmem_lane + (gi % ZF_GROUPS) * 96u + (uint32_t)((1 - dz) * 32) + c0, a[dz + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) a[dz + 1][j] = 0u;
        }
      }
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j)
        v[j] = (__uint_as_float(a[0][j]) + __uint_as_float(a[1][j])) + __uint_as_float(a[2][j]);
    };
    // same release rule as conv_zfold.cuh, signalled to the leader's barrier
    auto release_plane = [&](uint32_t gi0, int po) {
      __syncwarp();
      if (lane == 0) {
        if (po - 1 >= zin0) {
          const uint32_t gi = gi0 + (uint32_t)(po - 1 - zin0);
          mbar_arrive_leader(tempty0 + (gi % ZF_GROUPS) * 8u);
        }
        if (po == zend - 1) {
          const uint32_t gi = gi0 + (uint32_t)(po - zin0);
          mbar_arrive_leader(tempty0 + (gi % ZF_GROUPS) * 8u);
          if (po + 1 < zin1) mbar_arrive_leader(tempty0 + ((gi + 1u) % ZF_GROUPS) * 8u);
        }
      }
    };

    uint32_t gi0 = 0;
    if constexpr (EPI == EPI_STORE) {
      const int n0 = half * 32 + hs * 16;
      float bias[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) bias[j] = __ldg(p.bias + n0 + j);
      const size_t plane_elems = (size_t)p.H * p.W * p.out_cstride;
      const bool pool = p.pool_out != nullptr;

      auto do_plane = [&](int po, __nv_bfloat16* dst, bool in_xy, uint32_t (&pk)[8]) {
        float v[16];
        wait_plane(gi0, po);
        gather16(gi0, po, col_off, v);
        tc_fence_before();
        release_plane(gi0, po);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = leaky_relu(v[2 * j] + bias[2 * j]);
          const float c = leaky_relu(v[2 * j + 1] + bias[2 * j + 1]);
          pk[j] = pack_bf16x2(a, c);
        }
        if (in_xy) st_global_256(dst, pk);
      };

      for (int rd = 0; rd < my_rounds; ++rd) {
        const int tile = 2 * (pair_in_class + rd * pairs_per_class) + (int)rank;
        const bool live = tile < p.tiles_total;
        const int b = live ? tile / tiles_per_b : 0;
        const int r = live ? tile - b * tiles_per_b : 0;
        const int ty = r / p.ntx, tx = r - ty * p.ntx;
        const int x = p.ox + tx * 8 + rx, y = p.oy + ty * 16 + ry;
        const bool in_xy = live && x < p.W && y < p.H;
        __nv_bfloat16* dst = p.out + ((((size_t)b * p.D + p.oz) * p.H + y) * p.W + x) * p.out_cstride +
                             p.out_coff + n0;
        if (!pool) {
          for (int po = p.oz; po < zend; ++po) {
            uint32_t pk[8];
            do_plane(po, dst, in_xy, pk);
            dst += plane_elems;
          }
        } else {
          __nv_bfloat16* pdst =
              p.pool_out +
              ((((size_t)b * (p.D >> 1) + (p.oz >> 1)) * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1)) *
                  p.pool_cstride +
              p.pool_coff + n0;
          const size_t pplane = (size_t)(p.H >> 1) * (p.W >> 1) * p.pool_cstride;
          for (int po = p.oz; po < zend; po += 2) {
            uint32_t pa[8], pb[8];
            do_plane(po, dst, in_xy, pa);
            dst += plane_elems;
            do_plane(po + 1, dst, in_xy, pb);
            dst += plane_elems;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              __nv_bfloat162 m = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&pa[j]),
                                         *reinterpret_cast<__nv_bfloat162*>(&pb[j]));
              uint32_t mu = *reinterpret_cast<uint32_t*>(&m);
              uint32_t o1 = __shfl_xor_sync(0xffffffffu, mu, 1);
              m = __hmax2(m, *reinterpret_cast<__nv_bfloat162*>(&o1));
              mu = *reinterpret_cast<uint32_t*>(&m);
              uint32_t o8 = __shfl_xor_sync(0xffffffffu, mu, 8);
              m = __hmax2(m, *reinterpret_cast<__nv_bfloat162*>(&o8));
              pa[j] = *reinterpret_cast<uint32_t*>(&m);
            }
            if (in_xy && (lane & 9) == 0) st_global_256(pdst, pa);
            pdst += pplane;
          }
        }
        gi0 += (uint32_t)nin;
      }
    } else {
      float bias[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) bias[j] = __ldg(p.bias + half * 32 + j);
      const int t = p.trim;
      const int Dz = p.D - 2 * t, Hy = p.H - 2 * t, Wx = p.W - 2 * t;
      const size_t cstride = (size_t)Dz * Hy * Wx;
      for (int rd = 0; rd < my_rounds; ++rd) {
        const int tile = 2 * (pair_in_class + rd * pairs_per_class) + (int)rank;
        const bool live = tile < p.tiles_total;
        const int b = live ? tile / tiles_per_b : 0;
        const int r = live ? tile - b * tiles_per_b : 0;
        const int ty = r / p.ntx, tx = r - ty * p.ntx;
        const int x = p.ox + tx * 8 + rx, y = p.oy + ty * 16 + ry;
        const bool keep_xy = live && x >= t && x < p.W - t && y >= t && y < p.H - t;
        for (int po = p.oz; po < zend; ++po) {
          if (((po - p.oz) & 1) != hs) {  // the other warp set's plane
            release_plane(gi0, po);
            continue;
          }
          wait_plane(gi0, po);
          // the partials of planes po-1 and po are issued together (64 registers in flight), the
          // third follows while they are being added
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = bias[j];
          {
            uint32_t a0[32], a1[32];
            const bool has_prev = po - 1 >= zin0, has_next = po + 1 < zin1;
            const uint32_t gi = gi0 + (uint32_t)(po - zin0);
            if (has_prev) tmem_ld_32x32(tmem_lane + ((gi - 1u) % ZF_GROUPS) * 96u + 64u, a0);
            tmem_ld_32x32(tmem_lane + (gi % ZF_GROUPS) * 96u + 32u, a1);
            tmem_ld_wait();
            if (has_prev) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(a0[j]);
            }
            if (has_next) tmem_ld_32x32(tmem_lane + ((gi + 1u) % ZF_GROUPS) * 96u, a0);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(a1[j]);
            if (has_next) {
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(a0[j]);
            }
          }
          tc_fence_before();
          release_plane(gi0, po);
          if (keep_xy && po >= t && po < p.D - t) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = leaky_relu(v[j]);
            float* o = p.head_out + (size_t)b * p.head_c * cstride +
                       ((size_t)(po - t) * Hy + (y - t)) * Wx + (x - t);
#pragma unroll
            for (int oc = 0; oc < 8; ++oc) {
              if (oc < p.head_c) {  // uniform; weights are immediate constant-bank operands
                // four independent partial sums: the epilogue is latency-bound (2-3 warps per
                // scheduler), a 32-deep dependent FMA chain per channel would dominate it
                float r0 = p.head_b[oc], r1 = 0.f, r2 = 0.f, r3 = 0.f;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  r0 = fmaf(p.head_w[oc][j], v[j], r0);
                  r1 = fmaf(p.head_w[oc][j + 1], v[j + 1], r1);
                  r2 = fmaf(p.head_w[oc][j + 2], v[j + 2], r2);
                  r3 = fmaf(p.head_w[oc][j + 3], v[j + 3], r3);
                }
                float r = (r0 + r1) + (r2 + r3);
                if (p.apply_sigmoid) r = __fdividef(1.f, 1.f + __expf(-r));
                o[(size_t)oc * cstride] = r;
              }
            }
          }
        }
        gi0 += (uint32_t)nin;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA may free TMEM / exit while the pair's MMAs or signals are in flight
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512)
                 : "memory");
  }
}

}  // namespace exa
