// K1: 3x3x3 Conv3d (+ folded BatchNorm bias + LeakyReLU) as an implicit GEMM on the
// 5th-gen tensor cores: TMA -> shared memory -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM)
// -> tcgen05.ld -> fused epilogue.
//
// Replaces, per layer, reference unet3d.py:143-148 (Conv3d k=3 p=1 -> BatchNorm3d(eval)
// -> LeakyReLU(0.01)); with EPI_HEAD it also replaces unet3d.py:318 (1x1x1 OutConv),
// inference.py:158 (sigmoid) and inference.py:161-162 (trim).
//
// GEMM view: M = voxels (128-row tiles, one tile = tw x th x td x tb box of the NDHWC
// activation), N = Cout tile, K = 27 taps x Cin.  For every (tap, Cin chunk) the producer
// issues one 5-D TMA box load of the *shifted* voxel box -- out-of-bounds coordinates are
// zero-filled by TMA, which is exactly the conv's zero padding at the patch border -- and one
// 3-D TMA load of the weight slice [tap][n0:n0+N][c:c+KC].
#pragma once

#include "common.cuh"

namespace exa {

enum ConvEpilogue { EPI_STORE = 0, EPI_HEAD = 1 };

struct ConvArgs {
  // activation geometry (voxels) and tiling
  int B, D, H, W;
  int Cin, Cout;
  int tw, th, td, tb;       // tile box, tw*th*td*tb == 128
  int ntx, nty, ntz, ntb;   // tiles per dim
  int n_tiles_n;            // Cout / N
  int num_tiles;            // ntx*nty*ntz*ntb*n_tiles_n
  const float* bias;        // [Cout] folded BN bias
  // EPI_STORE: bf16 NDHWC output, channel stride/offset allow writing into a concat buffer
  __nv_bfloat16* out;
  int out_cstride;
  int out_coff;
  // EPI_HEAD: fused 1x1x1 head (+ optional sigmoid) on the 32 activated channels
  const float* head_w;      // [head_c][32]
  const float* head_b;      // [head_c]
  float* head_out;          // [B][head_c][D-2t][H-2t][W-2t] fp32
  int head_c;
  int trim;
  int apply_sigmoid;
};

// MT = number of 128-row M tiles per pipeline stage that share one B tile (MT = 2 halves the
// weight traffic per MMA; the kernel is bound by L2 -> shared-memory traffic, not by the MMAs)
template <int N, int KC, int MT = 1>
struct ConvSmem {
  static constexpr int A_BYTES = MT * 128 * KC * 2;
  static constexpr int B_BYTES = N * KC * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BUDGET = 200 * 1024;
  static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int BAR_BYTES = 256;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment
  static_assert(STAGES >= 3, "pipeline too shallow");
};

template <int N, int KC, int EPI, int MT = 1>
__global__ void __launch_bounds__(256, 1)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap tmap_x,
                    const __grid_constant__ CUtensorMap tmap_w, const ConvArgs p) {
  using S = ConvSmem<N, KC, MT>;
  constexpr int STAGES = S::STAGES;
  constexpr int ROW_BYTES = KC * 2;
  constexpr uint32_t TMEM_COLS = (2 * MT * N < 32) ? 32 : 2 * MT * N;  // two accumulator sets
  static_assert(2 * MT * N <= 512, "accumulators exceed TMEM");
  static_assert(EPI == EPI_STORE || N == 32, "fused head needs all 32 channels in one tile");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + STAGES * S::STAGE_BYTES);
  uint64_t* full_bar = bars;                  // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;        // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]       MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]    epilogue -> MMA
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 4);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), 4);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int kchunks = p.Cin / KC;
  const int kiters = 27 * kchunks;
  const int tiles_m_per_b = p.ntx * p.nty * p.ntz;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles_n;
        int m = tile / p.n_tiles_n;
        const int bt = m / tiles_m_per_b;
        m -= bt * tiles_m_per_b;
        const int zt = m / (p.ntx * p.nty);
        m -= zt * (p.ntx * p.nty);
        const int yt = m / p.ntx;
        const int xt = m - yt * p.ntx;
        const int x0 = xt * p.tw, y0 = yt * p.th, z0 = zt * p.td, b0 = bt * p.tb, n0 = nt * N;
        for (int tap = 0; tap < 27; ++tap) {
          const int kz = tap / 9, ky = (tap / 3) % 3, kx = tap % 3;
          for (int c = 0; c < kchunks; ++c) {
            mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1u);
            const uint32_t fb = smem_u32(&full_bar[stage]);
            mbar_expect_tx(fb, (uint32_t)S::STAGE_BYTES);
            const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
            tma_load_5d(sa, &tmap_x, fb, c * KC, x0 + kx - 1, y0 + ky - 1, z0 + kz - 1, b0);
            tma_load_3d(sa + S::A_BYTES, &tmap_w, fb, c * KC, n0, tap);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      // A barrier probe whose result is consumed at once costs ~150 cycles of tensor time
      // (tools/umma_probe.cu), more than half of a 4-MMA stage: every probe is issued one stage
      // (one tile) ahead and consumed later; the blocking wait only runs when the probe failed.
      bool ftok = false, ttok = false;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        if (!ttok) mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1u);
        ttok = mbar_test_wait(smem_u32(&tempty_bar[acc ^ 1]), ((acc ^ 1) == 0 ? acc_phase ^ 1u : acc_phase) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * MT * N);
        for (int it = 0; it < kiters; ++it) {
          if (!ftok) mbar_wait(smem_u32(&full_bar[stage]), phase);
          {
            const int ns = stage + 1 == STAGES ? 0 : stage + 1;
            ftok = mbar_test_wait(smem_u32(&full_bar[ns]), stage + 1 == STAGES ? phase ^ 1u : phase);
          }
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint64_t adesc = umma_smem_desc<ROW_BYTES>(sa);
          const uint64_t bdesc = umma_smem_desc<ROW_BYTES>(sa + S::A_BYTES);
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) {
#pragma unroll
            for (int m = 0; m < MT; ++m)
              umma_bf16(d_tmem + (uint32_t)(m * N), adesc + (uint64_t)((m * 128 * ROW_BYTES) >> 4) + 2 * k,
                        bdesc + 2 * k, idesc, (it | k) != 0);
          }
          umma_commit(smem_u32(&empty_bar[stage]));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(smem_u32(&tfull_bar[acc]));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue (4 warps, one TMEM lane quarter each) =====================
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles_n;
      int mt = tile / p.n_tiles_n;
      const int bt = mt / tiles_m_per_b;
      mt -= bt * tiles_m_per_b;
      const int zt = mt / (p.ntx * p.nty);
      mt -= zt * (p.ntx * p.nty);
      const int yt = mt / p.ntx;
      const int xt = mt - yt * p.ntx;
      const int n0 = nt * N;

      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const int row = m * 128 + q * 32 + lane;  // voxel of the tile box, x fastest
        const int r_tx = row % p.tw;
        const int r_ty = (row / p.tw) % p.th;
        const int r_tz = (row / (p.tw * p.th)) % p.td;
        const int r_tb = row / (p.tw * p.th * p.td);
        const int x = xt * p.tw + r_tx, y = yt * p.th + r_ty, z = zt * p.td + r_tz;
        const int b = bt * p.tb + r_tb;
        const bool valid = (x < p.W) && (y < p.H) && (z < p.D) && (b < p.B);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MT + m) * N);
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + (uint32_t)c0, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = leaky_relu(__uint_as_float(r[j]) + __ldg(p.bias + n0 + c0 + j));
          }
          if constexpr (EPI == EPI_STORE) {
            if (valid) {
              const size_t vox = (((size_t)b * p.D + z) * p.H + y) * p.W + x;
              uint4* dst =
                  reinterpret_cast<uint4*>(p.out + vox * p.out_cstride + p.out_coff + n0 + c0);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                uint4 o;
                o.x = pack_bf16x2(v[8 * g + 0], v[8 * g + 1]);
                o.y = pack_bf16x2(v[8 * g + 2], v[8 * g + 3]);
                o.z = pack_bf16x2(v[8 * g + 4], v[8 * g + 5]);
                o.w = pack_bf16x2(v[8 * g + 6], v[8 * g + 7]);
                dst[g] = o;
              }
            }
          } else {
            const int t = p.trim;
            const int Dz = p.D - 2 * t, Hy = p.H - 2 * t, Wx = p.W - 2 * t;
            const bool keep = valid && x >= t && x < p.W - t && y >= t && y < p.H - t && z >= t &&
                              z < p.D - t;
            if (keep) {
              for (int oc = 0; oc < p.head_c; ++oc) {
                float s = __ldg(p.head_b + oc);
#pragma unroll
                for (int j = 0; j < 32; ++j) s = fmaf(__ldg(p.head_w + oc * 32 + j), v[j], s);
                if (p.apply_sigmoid) s = 1.f / (1.f + expf(-s));
                const size_t o =
                    ((((size_t)b * p.head_c + oc) * Dz + (z - t)) * Hy + (y - t)) * Wx + (x - t);
                p.head_out[o] = s;
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace exa
