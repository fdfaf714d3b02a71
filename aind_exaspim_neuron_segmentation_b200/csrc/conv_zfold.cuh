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
//  * The single issuing thread must spend far fewer cycles per plane on bookkeeping than the
//    MMAs take and must never consume a barrier probe right after issuing it (a probe costs
//    ~150 cycles of tensor time, tools/umma_probe.cu).  Hence: (a) the steady-state planes run
//    one of three fully unrolled code variants selected by the ring position (no ring wrap /
//    wrap after two slots / wrap after one slot) whose descriptors are `base + immediate`;
//    (b) the issuer waits on a single barrier per plane (TMA data), probed one plane ahead; the
//    "TMEM slot drained" hand-shake is taken over by the TMA producer, which only loads an
//    input plane once the slots it will first touch are free; (c) eight epilogue warps (two per
//    TMEM lane quarter) keep the epilogue below the MMA time per plane.
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
  float head_w[8][32];       // EPI_HEAD: 1x1x1 head weights/bias as kernel parameters, so that
  float head_b[8];           // they are constant-bank operands of the FMAs (no loads)
  float* head_out;           // [B][head_c][D-2t][H-2t][W-2t]
  int head_c, trim, apply_sigmoid;
  int dbg;                   // EXA_ZF_DBG=8: record issuer cycles / wall time per CTA (results unchanged)
  long long* dbg_out;        // dbg & 8: [gridDim.x][4] issuer {cycles, ns, planes, 0}
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

constexpr int ZF_GROUPS = 5;     // TMEM accumulator groups of 3 x 32 fp32 columns (480 of 512)
constexpr int ZF_THREADS = 384;  // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 epilogue

// ---- UMMA shared-memory descriptors: (lo, hi) words; only the start-address field changes --
template <int ROWB>
__device__ __forceinline__ constexpr uint32_t zf_desc_hi(int sbo_rows) {
  return (uint32_t)((sbo_rows * ROWB) >> 4)      // SBO             [32,46)
         | (1u << 14)                            // version (sm_100) bit 46
         | ((ROWB == 128 ? 2u : (ROWB == 64 ? 4u : 6u)) << 29);  // swizzle mode [61,64): 128/64/32 B
}
__device__ __forceinline__ uint32_t zf_desc_lo(uint32_t smem_addr) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);  // start address [0,14), LBO (ignored)
}
__device__ __forceinline__ uint64_t zf_join(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

// All MMAs of one input plane: 9 taps x CIN/16 k-steps, each M=128, N=96, into the plane's own
// accumulator group (columns d0 .. d0+95 = partial sums of output planes z-1, z, z+1).  The first
// MMA overwrites, so groups need no zeroing and no plane is special.
template <int CIN>
__device__ __forceinline__ void zf_issue_plane(uint32_t d0, uint64_t a_desc, uint64_t w_desc) {
  using S = ZfSmem<CIN>;
  constexpr int ROWB = S::ROWB;
  constexpr int KSTEPS = CIN / 16;
  constexpr uint32_t I96 = umma_idesc_bf16(128, 96);
#pragma unroll
  for (int t = 0; t < 9; ++t) {
#pragma unroll
    for (int k = 0; k < KSTEPS; ++k) {
      // tap (ky,kx) shifts the halo view's start by ky*10+kx rows; the start-address field never
      // carries out of its 14 bits, so plain 64-bit adds of small immediates are exact
      const uint64_t ad = a_desc + (uint64_t)((((t / 3) * 10 + (t % 3)) * ROWB + k * 32) >> 4);
      const uint64_t bd = w_desc + (uint64_t)((t * S::W_TAP + k * 32) >> 4);
      umma_bf16(d0, ad, bd, I96, (t == 0 && k == 0) ? 0u : 1u);
    }
  }
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// one 32-byte (full sector) store
__device__ __forceinline__ void st_global_256(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

template <int CIN, int EPI>
__global__ void __launch_bounds__(ZF_THREADS, 1)
conv3x3_zfold_kernel(const __grid_constant__ CUtensorMap tmap_x,
                     const __grid_constant__ CUtensorMap tmap_w, const ZfArgs p) {
  using S = ZfSmem<CIN>;
  constexpr int STAGES = S::STAGES;
  // every epilogue warp releases every accumulator group once per use
  constexpr uint32_t TEMPTY_COUNT = 8;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + S::W_BYTES;
  uint64_t* bars = (uint64_t*)(smem_a + STAGES * S::A_STAGE);
  uint64_t* full_bar = bars;                             // [STAGES] TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;                   // [STAGES] MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;               // [5] MMA -> epilogue (group complete)
  uint64_t* tempty_bar = bars + 2 * STAGES + ZF_GROUPS;  // [5] epilogue -> MMA (group drained)
  uint64_t* w_bar = bars + 2 * STAGES + 2 * ZF_GROUPS;   // weights resident
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
    for (int s = 0; s < ZF_GROUPS; ++s) {
      mbar_init(smem_u32(&tfull_bar[s]), 1);
      mbar_init(smem_u32(&tempty_bar[s]), TEMPTY_COUNT);
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
  const int zin0 = max(p.oz - 1, 0);               // input planes [zin0, zin1) feed the
  const int zin1 = min(p.oz + p.nzp + 1, p.D);     // output planes [oz, oz + nzp)
  const int nin = zin1 - zin0;
  const int zend = p.oz + p.nzp;
  const int my_tiles = cta_in_class < p.tiles_total
                           ? (p.tiles_total - cta_in_class + ctas_per_class - 1) / ctas_per_class
                           : 0;
  // Input planes of this CTA are numbered consecutively across its tiles: gi = tile * nin +
  // (z - zin0).  Plane gi uses shared-memory stage gi % STAGES and accumulator group gi % 5.

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
    // Every input plane is the same straight-line sequence.  Two disjoint sets of loop-carried
    // variables: the descriptor / TMEM column / commit-address chain (`a_desc`, `ea`, `d0`, `tfa`:
    // add + wrap) and the mbarrier probe chains (`fa`, `fph`: TMA data; `ta`, `tph`: accumulator
    // group drained); each probe result is consumed one plane later.
    if (elect_one()) {
      mbar_wait(smem_u32(w_bar), 0);
      tc_fence_after();
      constexpr int ROWB = S::ROWB;
      // halo view: 8-row groups are 10 rows apart (A); weights: plain 8-row groups (B)
      const uint64_t w_desc_c = zf_join(zf_desc_lo(smem_u32(smem_w)), zf_desc_hi<ROWB>(8));
      const uint64_t a_desc0 = zf_join(zf_desc_lo(smem_u32(smem_a)), zf_desc_hi<ROWB>(10));
      const uint64_t a_desc_end = a_desc0 + (uint64_t)STAGES * (uint64_t)(S::A_STAGE >> 4);
      const uint32_t bar0 = smem_u32(bars);
      const uint32_t empty0 = bar0 + (uint32_t)STAGES * 8u;
      const uint32_t tfull0 = bar0 + 2u * (uint32_t)STAGES * 8u;
      const uint32_t tempty0 = tfull0 + (uint32_t)ZF_GROUPS * 8u;
      const uint32_t tempty_end = tempty0 + (uint32_t)ZF_GROUPS * 8u;
      uint64_t a_desc = a_desc0;     // descriptor of the current A stage
      uint32_t ea = empty0;          // empty_bar of the current stage
      uint32_t d0 = tmem_base;       // first column of the current accumulator group
      uint32_t tfa = tfull0;         // tfull_bar of the current group
      uint32_t fa = bar0, fph = 0;   // probe chain: full_bar of the current stage
      uint32_t ta = tempty0, tph = 1;  // probe chain: tempty_bar of the current group (first
                                       // round: parity 1 passes on a fresh barrier)
      bool ftok = false, ttok = false;  // already seen complete (probed one plane ahead)

      long long dbg_c0 = 0, dbg_t0 = 0;
      if (p.dbg & 8) {
        dbg_c0 = clock64();
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(dbg_t0));
      }
      const int total = my_tiles * nin;
      for (int it = 0; it < total; ++it) {
        if (!ftok) mbar_wait(fa, fph);
        if (!ttok) mbar_wait(ta, tph);
        // probe the next plane's barriers now, consume the answers one plane later
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
        // keep the 18..36 weight descriptors `base + immediate` instead of letting the compiler
        // hoist them into dozens of loop-invariant registers
        uint64_t w_desc = w_desc_c;
        asm volatile("" : "+l"(w_desc));
        zf_issue_plane<CIN>(d0, a_desc, w_desc);
        umma_commit(ea);   // shared-memory stage consumed once these MMAs retire
        umma_commit(tfa);  // accumulator group complete
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
    // ===================== epilogue: 8 warps, two per TMEM lane quarter =====================
    // Output plane po = sum of up to three partials: columns [64,96) of input plane po-1's group,
    // [32,64) of plane po's and [0,32) of plane po+1's.
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int hs = (warp - 4) >> 2;  // 0/1: column half (EPI_STORE) or plane parity (EPI_HEAD)
    const int row = q * 32 + lane;
    const int rx = row & 7, ry = row >> 3;
    const uint32_t tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar);
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    constexpr int NC = EPI == EPI_STORE ? 16 : 32;    // accumulator columns per thread
    const uint32_t col_off = EPI == EPI_STORE ? (uint32_t)(hs * 16) : 0u;

    // wait for the partials of output plane po and sum them into v[]  (gi0 = number of this
    // tile's first input plane)
    auto gather_plane = [&](uint32_t gi0, int po, float (&v)[NC]) {
      const int zl = po + 1 < zin1 ? po + 1 : po;  // latest contributing input plane
      {
        const uint32_t gi = gi0 + (uint32_t)(zl - zin0);
        mbar_wait(tfull0 + (gi % ZF_GROUPS) * 8u, (gi / ZF_GROUPS) & 1u);
        tc_fence_after();
      }
#pragma unroll
      for (int j = 0; j < NC; ++j) v[j] = 0.f;
      {
#pragma unroll
        for (int dz = -1; dz <= 1; ++dz) {
          const int z = po + dz;
          if (z >= zin0 && z < zin1) {
            const uint32_t gi = gi0 + (uint32_t)(z - zin0);
            const uint32_t taddr =
                tmem_lane + (gi % ZF_GROUPS) * 96u + (uint32_t)((1 - dz) * 32) + col_off;
            uint32_t acc[NC];
            if constexpr (NC == 16) tmem_ld_32x16(taddr, acc);
            else tmem_ld_32x32(taddr, acc);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < NC; ++j) v[j] += __uint_as_float(acc[j]);
          }
        }
      }
      tc_fence_before();
    };
    // Release rule, executed by EVERY epilogue warp for EVERY output plane po in order (also for
    // planes whose partials the warp does not read): after plane po this warp will not read the
    // group of input plane po-1 again; after the column's last plane nor those of planes po, po+1.
    auto release_plane = [&](uint32_t gi0, int po) {
      __syncwarp();
      if (lane == 0) {
        if (po - 1 >= zin0) {
          const uint32_t gi = gi0 + (uint32_t)(po - 1 - zin0);
          mbar_arrive(tempty0 + (gi % ZF_GROUPS) * 8u);
        }
        if (po == zend - 1) {
          const uint32_t gi = gi0 + (uint32_t)(po - zin0);
          mbar_arrive(tempty0 + (gi % ZF_GROUPS) * 8u);
          if (po + 1 < zin1) mbar_arrive(tempty0 + ((gi + 1u) % ZF_GROUPS) * 8u);
        }
      }
    };

    uint32_t gi0 = 0;  // number of this tile's first input plane
    if constexpr (EPI == EPI_STORE) {
      const int n0 = half * 32 + hs * 16;
      float bias[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) bias[j] = __ldg(p.bias + n0 + j);
      const size_t plane_elems = (size_t)p.H * p.W * p.out_cstride;
      const bool pool = p.pool_out != nullptr;

      // one plane: partial sums -> bias + LeakyReLU -> bf16 -> one 32 B store per voxel
      auto do_plane = [&](int po, __nv_bfloat16* dst, bool in_xy, uint32_t (&pk)[8]) {
        float v[16];
        gather_plane(gi0, po, v);
        release_plane(gi0, po);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = leaky_relu(v[2 * j] + bias[2 * j]);
          const float c = leaky_relu(v[2 * j + 1] + bias[2 * j + 1]);
          pk[j] = pack_bf16x2(a, c);
        }
        if (in_xy) st_global_256(dst, pk);
      };

      for (int tile = cta_in_class; tile < p.tiles_total; tile += ctas_per_class) {
        const int b = tile / tiles_per_b;
        const int r = tile - b * tiles_per_b;
        const int ty = r / p.ntx, tx = r - ty * p.ntx;
        const int x = p.ox + tx * 8 + rx, y = p.oy + ty * 16 + ry;
        const bool in_xy = x < p.W && y < p.H;
        __nv_bfloat16* dst = p.out + ((((size_t)b * p.D + p.oz) * p.H + y) * p.W + x) * p.out_cstride +
                             p.out_coff + n0;
        if (!pool) {
          for (int po = p.oz; po < zend; ++po) {
            uint32_t pk[8];
            do_plane(po, dst, in_xy, pk);
            dst += plane_elems;
          }
        } else {
          // fused MaxPool3d(2): z pair in this thread, x pair = lane ^ 1, y pair = lane ^ 8; max
          // commutes with the bf16 rounding, so this equals pooling the stored tensor.
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
      // fused 1x1x1 head (+sigmoid) and trim: every thread needs all 32 channels of its voxel,
      // so the two warp sets take alternate planes instead of column halves
      float bias[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) bias[j] = __ldg(p.bias + half * 32 + j);
      const int t = p.trim;
      const int Dz = p.D - 2 * t, Hy = p.H - 2 * t, Wx = p.W - 2 * t;
      const size_t cstride = (size_t)Dz * Hy * Wx;
      for (int tile = cta_in_class; tile < p.tiles_total; tile += ctas_per_class) {
        const int b = tile / tiles_per_b;
        const int r = tile - b * tiles_per_b;
        const int ty = r / p.ntx, tx = r - ty * p.ntx;
        const int x = p.ox + tx * 8 + rx, y = p.oy + ty * 16 + ry;
        const bool keep_xy = x >= t && x < p.W - t && y >= t && y < p.H - t;
        for (int po = p.oz; po < zend; ++po) {
          if (((po - p.oz) & 1) != hs) {  // the other warp set's plane
            release_plane(gi0, po);
            continue;
          }
          float v[32];
          gather_plane(gi0, po, v);
          release_plane(gi0, po);
          if (keep_xy && po >= t && po < p.D - t) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = leaky_relu(v[j] + bias[j]);
            float* o = p.head_out + (size_t)b * p.head_c * cstride +
                       ((size_t)(po - t) * Hy + (y - t)) * Wx + (x - t);
#pragma unroll
            for (int oc = 0; oc < 8; ++oc) {
              if (oc < p.head_c) {  // uniform; weights are immediate constant-bank operands
                float s = p.head_b[oc];
#pragma unroll
                for (int j = 0; j < 32; ++j) s = fmaf(p.head_w[oc][j], v[j], s);
                if (p.apply_sigmoid) s = __fdividef(1.f, 1.f + __expf(-s));
                o[(size_t)oc * cstride] = s;
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
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace exa
