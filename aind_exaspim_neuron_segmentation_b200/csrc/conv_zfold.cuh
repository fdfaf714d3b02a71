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
  const float* head_w;       // EPI_HEAD: [head_c][32]
  const float* head_b;       // [head_c]
  float* head_out;           // [B][head_c][D-2t][H-2t][W-2t]
  int head_c, trim, apply_sigmoid;
  int dbg;                   // development only (EXA_ZF_DBG): timing experiments, wrong results
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

constexpr int ZF_RING = 16;      // TMEM slots of 32 fp32 columns
constexpr int ZF_THREADS = 384;  // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4..11 epilogue

// ---- UMMA shared-memory descriptors split in (lo, hi) words: only `lo` ever changes ------
template <int ROWB>
__device__ __forceinline__ constexpr uint32_t zf_desc_hi(int sbo_rows) {
  return (uint32_t)((sbo_rows * ROWB) >> 4)      // SBO             [32,46)
         | (1u << 14)                            // version (sm_100) bit 46
         | ((ROWB == 128 ? 2u : 4u) << 29);      // swizzle mode    [61,64)
}
__device__ __forceinline__ uint32_t zf_desc_lo(uint32_t smem_addr) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);  // start address [0,14), LBO (ignored)
}
__device__ __forceinline__ uint64_t zf_join(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

// All MMAs of one steady-state input plane (planes zi-1, zi, zi+1 all valid; zi+1 is touched
// for the first time).  w = ring slot of plane zi-1.
//   V == 0 (w <= 13): one N=96 group at column 32*w;
//   V == 1 (w == 14): slots (14, 15 | 0);   V == 2 (w == 15): slots (15 | 0, 1).
template <int CIN, int V>
__device__ __forceinline__ void zf_issue_steady(uint32_t tmem_base, uint32_t w, uint32_t a_lo,
                                                uint32_t w_lo) {
  using S = ZfSmem<CIN>;
  constexpr int ROWB = S::ROWB;
  constexpr int KSTEPS = CIN / 16;
  constexpr uint32_t A_HI = zf_desc_hi<ROWB>(10);  // halo view: 8-row groups are 10 rows apart
  constexpr uint32_t B_HI = zf_desc_hi<ROWB>(8);
  constexpr uint32_t I32 = umma_idesc_bf16(128, 32), I64 = umma_idesc_bf16(128, 64),
                     I96 = umma_idesc_bf16(128, 96);
  constexpr uint32_t BROW = (32 * ROWB) >> 4;  // descriptor units per 32 B rows (one kz block)
  const uint32_t d0 = tmem_base + (V == 0 ? w * 32u : (V == 1 ? 448u : 480u));
#pragma unroll
  for (int t = 0; t < 9; ++t) {
#pragma unroll
    for (int k = 0; k < KSTEPS; ++k) {
      const bool first = (t == 0 && k == 0);
      // tap (ky,kx) shifts the halo view's start by ky*10+kx rows
      const uint64_t ad = zf_join(a_lo + (uint32_t)((((t / 3) * 10 + (t % 3)) * ROWB + k * 32) >> 4), A_HI);
      const uint32_t b = w_lo + (uint32_t)((t * S::W_TAP + k * 32) >> 4);
      if (V == 0) {
        if (first) {
          umma_bf16(d0, ad, zf_join(b, B_HI), I64, 1u);
          umma_bf16(d0 + 64u, ad, zf_join(b + 2 * BROW, B_HI), I32, 0u);
        } else {
          umma_bf16(d0, ad, zf_join(b, B_HI), I96, 1u);
        }
      } else if (V == 1) {
        umma_bf16(d0, ad, zf_join(b, B_HI), I64, 1u);
        umma_bf16(tmem_base, ad, zf_join(b + 2 * BROW, B_HI), I32, first ? 0u : 1u);
      } else {
        umma_bf16(d0, ad, zf_join(b, B_HI), I32, 1u);
        if (first) {
          umma_bf16(tmem_base, ad, zf_join(b + BROW, B_HI), I32, 1u);
          umma_bf16(tmem_base + 32u, ad, zf_join(b + 2 * BROW, B_HI), I32, 0u);
        } else {
          umma_bf16(tmem_base, ad, zf_join(b + BROW, B_HI), I64, 1u);
        }
      }
    }
  }
}

// Edge planes of a column (first/last input planes: some of zi-1, zi, zi+1 are outside the
// output range): one N=32 MMA per valid z tap.  Four planes per column, so speed is secondary.
template <int CIN>
__device__ __noinline__ void zf_issue_edge(uint32_t tmem_base, uint32_t a_lo, uint32_t w_lo,
                                           uint32_t col0, uint32_t col1, uint32_t col2,
                                           uint32_t flags) {
  using S = ZfSmem<CIN>;
  constexpr int ROWB = S::ROWB;
  constexpr int KSTEPS = CIN / 16;
  constexpr uint32_t A_HI = zf_desc_hi<ROWB>(10);
  constexpr uint32_t B_HI = zf_desc_hi<ROWB>(8);
  constexpr uint32_t I32 = umma_idesc_bf16(128, 32);
  constexpr uint32_t BROW = (32 * ROWB) >> 4;
  const uint32_t col[3] = {col0, col1, col2};
#pragma unroll 1
  for (int t = 0; t < 9; ++t) {
#pragma unroll
    for (int k = 0; k < KSTEPS; ++k) {
      const uint64_t ad = zf_join(a_lo + (uint32_t)((((t / 3) * 10 + (t % 3)) * ROWB + k * 32) >> 4), A_HI);
      const uint32_t b = w_lo + (uint32_t)((t * S::W_TAP + k * 32) >> 4);
#pragma unroll
      for (int kzr = 0; kzr < 3; ++kzr) {
        if (flags & (1u << kzr)) {  // valid
          const uint32_t acc = (t == 0 && k == 0 && (flags & (8u << kzr))) ? 0u : 1u;  // fresh
          umma_bf16(tmem_base + col[kzr], ad, zf_join(b + kzr * BROW, B_HI), I32, acc);
        }
      }
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
  // epilogue arrivals per TMEM slot: EPI_STORE splits the 32 columns over two warps per lane
  // quarter (8 warps touch every plane); EPI_HEAD alternates planes between the two warp sets
  constexpr uint32_t TEMPTY_COUNT = EPI == EPI_STORE ? 8 : 4;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem;
  uint8_t* smem_a = smem + S::W_BYTES;
  uint64_t* bars = (uint64_t*)(smem_a + STAGES * S::A_STAGE);
  uint64_t* full_bar = bars;                          // [STAGES] TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;                // [STAGES] MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;            // [16] MMA -> epilogue (slot complete)
  uint64_t* tempty_bar = bars + 2 * STAGES + ZF_RING; // [16] epilogue -> TMA producer (slot drained)
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
  const int zin0 = max(p.oz - 1, 0);
  const int zin1 = min(p.oz + p.nzp + 1, p.D);
  const int zend = p.oz + p.nzp;

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
      uint32_t gbase = 0;  // running count of output planes handled by this CTA
      for (int tile = cta_in_class; tile < p.tiles_total; tile += ctas_per_class) {
        const int b = tile / tiles_per_b;
        const int r = tile - b * tiles_per_b;
        const int ty = r / p.ntx, tx = r - ty * p.ntx;
        const int x0 = p.ox + tx * 8, y0 = p.oy + ty * 16;
        for (int zi = zin0; zi < zin1; ++zi) {
          // TMEM slots first touched by the MMAs of input plane zi (plane zi+1; at the first
          // input plane also plane zi) must have been drained by the epilogue.  Doing this
          // hand-shake here keeps it off the MMA issuer's critical path.
          if (!(p.dbg & 4)) {
            if (zi == zin0 && zi >= p.oz) {
              const uint32_t g = gbase + (uint32_t)(zi - p.oz);
              mbar_wait(smem_u32(&tempty_bar[g % ZF_RING]), ((g / ZF_RING) & 1u) ^ 1u);
            }
            if (zi + 1 >= p.oz && zi + 1 < zend) {
              const uint32_t g = gbase + (uint32_t)(zi + 1 - p.oz);
              mbar_wait(smem_u32(&tempty_bar[g % ZF_RING]), ((g / ZF_RING) & 1u) ^ 1u);
            }
          }
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_expect_tx(fb, (uint32_t)S::A_TX_BYTES);
          tma_load_5d(smem_u32(smem_a + stage * S::A_STAGE), &tmap_x, fb, 0, x0 - 1, y0 - 1, zi, b);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        gbase += (uint32_t)p.nzp;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // Two disjoint sets of loop-carried variables: the descriptor / TMEM column / commit-address
    // chain (`a_lo`, `ea`, `w`: add + wrap, nothing derived from the barrier variables) and the
    // mbarrier probe chain (`fa`, `phase`); the probe result is consumed one plane later.
    if (elect_one()) {
      mbar_wait(smem_u32(w_bar), 0);
      tc_fence_after();
      const uint32_t w_lo_c = zf_desc_lo(smem_u32(smem_w));
      const uint32_t a_lo0 = zf_desc_lo(smem_u32(smem_a));
      const uint32_t a_lo_end = a_lo0 + (uint32_t)STAGES * (uint32_t)(S::A_STAGE >> 4);
      const uint32_t bar0 = smem_u32(bars);
      const uint32_t fa_end = bar0 + (uint32_t)STAGES * 8u;        // == &empty_bar[0]
      const uint32_t ea_end = bar0 + 2u * (uint32_t)STAGES * 8u;   // == &tfull_bar[0]
      const uint32_t tfull0 = ea_end;
      const int my_tiles = cta_in_class < p.tiles_total
                               ? (p.tiles_total - cta_in_class + ctas_per_class - 1) / ctas_per_class
                               : 0;
      const int oz = p.oz, nzp = p.nzp;
      const int n_steady = nzp - 2;
      const bool nowait = (p.dbg & 2) != 0;
      uint32_t a_lo = a_lo0;    // descriptor word of the current A stage        (uniform chain)
      uint32_t ea = fa_end;     // empty_bar of the current stage                (uniform chain)
      uint32_t w = 15u;         // ring slot of plane zi-1                       (uniform chain)
      uint32_t fa = bar0;       // full_bar of the current stage                 (probe chain)
      uint32_t phase = 0;       //                                               (probe chain)
      bool tok = false;         // full_bar of the current stage already seen complete (peeked)

      // wait for the TMA data of the current stage; probe the next stage now and consume the
      // answer one plane later
      auto acquire = [&]() {
        if (!tok && !nowait) mbar_wait(fa, phase);
        fa += 8u;
        if (fa == fa_end) {
          fa = bar0;
          phase ^= 1u;
        }
        tok = mbar_test_wait(fa, phase);
        tc_fence_after();
      };
      auto next_stage = [&]() {
        umma_commit(ea);  // stage consumed once the MMAs issued so far retire
        ea += 8u;
        a_lo += (uint32_t)(S::A_STAGE >> 4);
        if (a_lo == a_lo_end) {
          a_lo = a_lo0;
          ea = fa_end;
        }
      };
      // first/last input planes of a column: some of zi-1, zi, zi+1 are not output planes
      auto edge_plane = [&](int zi) {
        acquire();
        uint32_t w_lo = w_lo_c;
        asm volatile("" : "+r"(w_lo));
        uint32_t flags = 0;
#pragma unroll
        for (int kzr = 0; kzr < 3; ++kzr) {
          const int po = zi - 1 + kzr;
          if (po >= oz && po < zend) {
            flags |= 1u << kzr;
            if (zi == max(po - 1, zin0)) flags |= 8u << kzr;  // first touch of this slot
          }
        }
        zf_issue_edge<CIN>(tmem_base, a_lo, w_lo, w * 32u, ((w + 1u) % ZF_RING) * 32u,
                           ((w + 2u) % ZF_RING) * 32u, flags);
        next_stage();
        // completed output planes: zi-1 always; zi too when it is the last input plane
        if (zi > oz && zi - 1 < zend) umma_commit(tfull0 + w * 8u);
        if (zi == zin1 - 1 && zi >= oz && zi < zend) umma_commit(tfull0 + ((w + 1u) % ZF_RING) * 8u);
        w = (w + 1u) % ZF_RING;
      };

      for (int tl = 0; tl < my_tiles; ++tl) {
        // ring slot of plane zin0-1: plane oz sits at slot (tl*nzp) % 16
        w = ((uint32_t)tl * (uint32_t)nzp + (uint32_t)(zin0 - 1 - oz)) % ZF_RING;
        int zi = zin0;
        for (; zi <= oz && zi < zin1; ++zi) edge_plane(zi);
        for (int i = 0; i < n_steady; ++i) {
          acquire();
          // keep the 18..36 weight descriptors `base + immediate` instead of letting the
          // compiler hoist them into dozens of loop-invariant registers
          uint32_t w_lo = w_lo_c;
          asm volatile("" : "+r"(w_lo));
          if (w <= 13u) {
            zf_issue_steady<CIN, 0>(tmem_base, w, a_lo, w_lo);
          } else if (w == 14u) {
            zf_issue_steady<CIN, 1>(tmem_base, w, a_lo, w_lo);
          } else {
            zf_issue_steady<CIN, 2>(tmem_base, w, a_lo, w_lo);
          }
          next_stage();
          umma_commit(tfull0 + w * 8u);  // plane zi-1 is complete
          w = (w + 1u) % ZF_RING;
        }
        zi += n_steady > 0 ? n_steady : 0;
        for (; zi < zin1; ++zi) edge_plane(zi);
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue: 8 warps, two per TMEM lane quarter =====================
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int hs = (warp - 4) >> 2;  // 0/1: column half (EPI_STORE) or plane parity (EPI_HEAD)
    const int row = q * 32 + lane;
    const int rx = row & 7, ry = row >> 3;
    const uint32_t tfull0 = smem_u32(tfull_bar), tempty0 = smem_u32(tempty_bar);
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool skip = (p.dbg & 1) != 0;
    uint32_t gbase = 0;

    if constexpr (EPI == EPI_STORE) {
      const int n0 = half * 32 + hs * 16;
      float bias[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) bias[j] = __ldg(p.bias + n0 + j);
      const size_t plane_elems = (size_t)p.H * p.W * p.out_cstride;
      const bool pool = p.pool_out != nullptr;

      // one plane: TMEM -> registers -> bias + LeakyReLU -> bf16 -> one 32 B store per voxel
      auto do_plane = [&](uint32_t g, __nv_bfloat16* dst, bool in_xy, uint32_t (&pk)[8]) {
        const uint32_t slot = g % ZF_RING;
        mbar_wait(tfull0 + slot * 8u, (g / ZF_RING) & 1u);
        tc_fence_after();
        if (skip) {
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty0 + slot * 8u);
#pragma unroll
          for (int j = 0; j < 8; ++j) pk[j] = 0;
          return;
        }
        uint32_t acc[16];
        tmem_ld_32x16(tmem_lane + slot * 32u + (uint32_t)(hs * 16), acc);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + slot * 8u);  // slot is in registers now
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = leaky_relu(__uint_as_float(acc[2 * j]) + bias[2 * j]);
          const float c = leaky_relu(__uint_as_float(acc[2 * j + 1]) + bias[2 * j + 1]);
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
          for (int i = 0; i < p.nzp; ++i) {
            uint32_t pk[8];
            do_plane(gbase + (uint32_t)i, dst, in_xy, pk);
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
          for (int i = 0; i < p.nzp; i += 2) {
            uint32_t pa[8], pb[8];
            do_plane(gbase + (uint32_t)i, dst, in_xy, pa);
            dst += plane_elems;
            do_plane(gbase + (uint32_t)i + 1u, dst, in_xy, pb);
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
        gbase += (uint32_t)p.nzp;
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
        for (int i = hs; i < p.nzp; i += 2) {
          const uint32_t g = gbase + (uint32_t)i;
          const uint32_t slot = g % ZF_RING;
          const int po = p.oz + i;
          mbar_wait(tfull0 + slot * 8u, (g / ZF_RING) & 1u);
          tc_fence_after();
          uint32_t acc[32];
          if (!skip) {
            tmem_ld_32x32(tmem_lane + slot * 32u, acc);
            tmem_ld_wait();
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty0 + slot * 8u);
          if (!skip && keep_xy && po >= t && po < p.D - t) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = leaky_relu(__uint_as_float(acc[j]) + bias[j]);
            float* o = p.head_out + (size_t)b * p.head_c * cstride +
                       ((size_t)(po - t) * Hy + (y - t)) * Wx + (x - t);
            for (int oc = 0; oc < p.head_c; ++oc) {
              float s = __ldg(p.head_b + oc);
#pragma unroll
              for (int j = 0; j < 32; ++j) s = fmaf(__ldg(p.head_w + oc * 32 + j), v[j], s);
              if (p.apply_sigmoid) s = 1.f / (1.f + expf(-s));
              o[(size_t)oc * cstride] = s;
            }
          }
        }
        gbase += (uint32_t)p.nzp;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace exa
