// Hardware probe (development tool, not on the product path).
//
// Answers three design questions for the implicit-GEMM conv kernel on a real
// B200 before the kernel is built around them:
//  1. descriptor sanity: TMA(SWIZZLE_128B / 64B) -> tcgen05.mma -> tcgen05.ld
//     reproduces a CPU GEMM exactly (small-integer data);
//  2. "halo view": can one TMA-loaded (18 x 10) halo tile be re-used for all
//     9 in-plane taps by only moving the UMMA descriptor start address by
//     (ky*10 + kx) rows with SBO = 10 rows (unaligned 8-row groups)?
//  3. issue-rate of M=128 MMAs from shared memory as a function of N.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_probe umma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../common.cuh"
#include "../tmap.h"

using namespace exa;

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);   \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

struct ProbeArgs {
  int halo;         // 1: halo box (10 x 18) + offset descriptor; 0: plain (8 x 16) box at shifted coords
  int ky, kx;       // in-plane tap
  int z;            // input plane
  int tap;          // weight tap index for B
  int n;            // MMA N
  int use_base_offset;
  int b_row_off;    // B descriptor starts at this row (multiple of 8)
  float* out;       // [128][n]
  int dcol;         // first accumulator column (TMEM wrap-around test when dcol + n > 512)
};

template <int ROW_BYTES>
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
             const ProbeArgs a) {
  constexpr int KC = ROW_BYTES / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                       // up to 180 rows
  uint8_t* sB = smem + 32768;               // up to 256 rows
  uint64_t* bars = (uint64_t*)(smem + 32768 + 32768);
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_init(smem_u32(&bars[1]), 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (threadIdx.x == 0) {
    const int rowsA = a.halo ? 180 : 128;
    mbar_expect_tx(smem_u32(&bars[0]), rowsA * ROW_BYTES + 256 * ROW_BYTES);
    if (a.halo)
      tma_load_5d(smem_u32(sA), &tmap_x, smem_u32(&bars[0]), 0, -1, -1, a.z, 0);
    else
      tma_load_5d(smem_u32(sA), &tmap_x, smem_u32(&bars[0]), 0, a.kx - 1, a.ky - 1, a.z, 0);
    tma_load_3d(smem_u32(sB), &tmap_w, smem_u32(&bars[0]), 0, 0, a.tap);
    mbar_wait(smem_u32(&bars[0]), 0);
    tc_fence_after();
    uint32_t a_addr = smem_u32(sA);
    uint64_t adesc;
    if (a.halo) {
      a_addr += (a.ky * 10 + a.kx) * ROW_BYTES;
      adesc = umma_smem_desc<ROW_BYTES>(a_addr);
      // SBO = 10 rows
      adesc &= ~((uint64_t)0x3FFF << 32);
      adesc |= (uint64_t)((10 * ROW_BYTES) >> 4) << 32;
      if (a.use_base_offset) adesc |= (uint64_t)((a_addr >> 7) & 7) << 49;
    } else {
      adesc = umma_smem_desc<ROW_BYTES>(a_addr);
    }
    uint64_t bdesc = umma_smem_desc<ROW_BYTES>(smem_u32(sB) + a.b_row_off * ROW_BYTES);
    const uint32_t idesc = umma_idesc_bf16(128, a.n);
    for (int k = 0; k < KC / 16; ++k) umma_bf16(tmem + (uint32_t)a.dcol, adesc + 2 * k, bdesc + 2 * k, idesc, k > 0);
    umma_commit(smem_u32(&bars[1]));
  }
  __syncwarp();
  mbar_wait(smem_u32(&bars[1]), 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < a.n; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)((a.dcol + c0) % 512), r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) a.out[row * a.n + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ---- issue-rate probe ------------------------------------------------------
// 32 MMAs fully unrolled per loop trip with compile-time descriptor offsets, so the
// single issuing thread is not the limiter.  ACCS = number of TMEM accumulators cycled.
template <int N, int ACCS>
__global__ void __launch_bounds__(128, 1)
rate_kernel(int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + 196608);
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 196608 / 16; i += blockDim.x) ((uint4*)smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp_u == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t a0 = umma_smem_desc<128>(smem_u32(smem));
    const uint64_t b0 = umma_smem_desc<128>(smem_u32(smem) + 98304);
    long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; it += 32) {
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          const uint64_t adesc = a0 + (uint64_t)(((u >> 2) & 3) * 1024 + 2 * (u & 3));
          const uint64_t bdesc = b0 + (uint64_t)(((u >> 2) & 1) * 2048 + 2 * (u & 3));
          umma_bf16(tmem + (uint32_t)(((u >> 2) % ACCS) * N), adesc, bdesc, idesc, 1);
        }
      }
      umma_commit(smem_u32(&bars[0]));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bars[0]), 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// Same, but A descriptors are the nine "halo views" of the z-folded conv kernel: start address
// shifted by (ky*10+kx) rows, 8-row groups 10 rows apart (unaligned swizzle atoms), or -- with
// ALIGNED -- the 3-copy layout (x-shifted copies with pitch 8: start ky*8 rows, SBO 8 rows).
template <int N, int ROWB, bool ALIGNED>
__global__ void __launch_bounds__(128, 1)
rate_halo_kernel(int iters, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + 196608);
  uint32_t* tmem_slot = (uint32_t*)(bars + 4);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 196608 / 16; i += blockDim.x) ((uint4*)smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bars[0]), 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if (warp_u == 0) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, N);
    uint64_t a0 = umma_smem_desc<ROWB>(smem_u32(smem));
    if (!ALIGNED) {
      a0 &= ~((uint64_t)0x3FFF << 32);
      a0 |= (uint64_t)((10 * ROWB) >> 4) << 32;
    }
    const uint64_t b0 = umma_smem_desc<ROWB>(smem_u32(smem) + 98304);
    long long t0 = clock64();
    if (elect_one()) {
      for (int it = 0; it < iters; it += 36) {
#pragma unroll
        for (int t = 0; t < 9; ++t) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int kk = k % (ROWB / 32);
            const int rows = ALIGNED ? ((t % 3) * 144 + (t / 3) * 8) : ((t / 3) * 10 + (t % 3));
            const uint64_t adesc = a0 + (uint64_t)((rows * ROWB + kk * 32) >> 4);
            const uint64_t bdesc = b0 + (uint64_t)(((t % 2) * N * ROWB + kk * 32) >> 4);  // stays inside smem
            umma_bf16(tmem, adesc, bdesc, idesc, 1);
          }
        }
      }
      umma_commit(smem_u32(&bars[0]));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bars[0]), 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int N, int ROWB, bool ALIGNED>
static void run_rate_halo(int sms, long long* dcyc) {
  const size_t smem_bytes = 196608 + 64 + 1024;
  CK(cudaFuncSetAttribute(rate_halo_kernel<N, ROWB, ALIGNED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  const int iters = 36 * 256;
  rate_halo_kernel<N, ROWB, ALIGNED><<<sms, 128, smem_bytes>>>(iters, dcyc);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("rate halo kernel error %s\n", cudaGetErrorString(e)); exit(3); }
  std::vector<long long> h(sms);
  CK(cudaMemcpy(h.data(), dcyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (auto v : h) mx = v > mx ? v : mx;
  double cyc = (double)mx / iters;
  printf("RATE-HALO N=%3d rowbytes=%3d %-9s : %.2f cyc/MMA (ideal %.1f) -> %.1f%% of tensor peak\n", N, ROWB,
         ALIGNED ? "aligned" : "unaligned", cyc, N / 2.0, 100.0 * (N / 2.0) / cyc);
}


// ---- issuer-loop emulation --------------------------------------------------------------
// One "plane" = 9 taps x KSTEPS halo-view MMAs (N=96) as in conv_zfold.cuh, plus the barrier
// traffic the real issuer thread has per plane.  Answers: what do tcgen05.commit and mbarrier
// probes cost when interleaved with the MMA stream of the single issuing thread?
//  MODE 0 MMAs only | 1 +1 commit | 2 +2 commits | 3 +2 commits +1 blocking try_wait (complete
//  barrier) | 4 +2 commits +2 try_waits | 5 +2 commits +2 test_wait peeks consumed next plane |
//  6 MODE 0 with the first k-step split in three N=32 MMAs | 7 MODE 2 + split first k-step
template <int ROWB, int MODE>
__global__ void __launch_bounds__(128, 1)
loop_kernel(int planes, long long* cycles, int fill) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + 196608);
  uint32_t* tmem_slot = (uint32_t*)(bars + 6);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 196608 / 16; i += blockDim.x) {
    uint4 v = make_uint4(0, 0, 0, 0);
    if (fill) {  // pseudo-random bf16 values in (-2, 2): realistic operand toggling
      uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
      auto nxt = [&]() { h = h * 1664525u + 1013904223u; return ((h >> 9) & 0x807F807Fu) | 0x3F003F00u; };
      v = make_uint4(nxt(), nxt(), nxt(), nxt());
    }
    ((uint4*)smem)[i] = v;
  }
  if (threadIdx.x == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_fence_init();
    mbar_arrive(smem_u32(&bars[3]));  // phase 0 of bars[3], bars[4] complete for good
    mbar_arrive(smem_u32(&bars[4]));
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  constexpr int KSTEPS = ROWB / 32;
  if (warp_u == 0) {
    uint64_t a0 = umma_smem_desc<ROWB>(smem_u32(smem));
    a0 &= ~((uint64_t)0x3FFF << 32);
    a0 |= (uint64_t)((10 * ROWB) >> 4) << 32;
    const uint64_t b0 = umma_smem_desc<ROWB>(smem_u32(smem) + (ROWB == 64 ? 98304 : 65536));
    long long t0 = clock64();
    if (elect_one()) {
      bool tok0 = true, tok1 = true;
      uint32_t wring = 0;
#pragma unroll 1
      for (int pl = 0; pl < planes; ++pl) {
        // MODE 8/9: the TMEM column operand is a loop-carried value in ONE register that is
        // rewritten every plane (MODE 9 also commits twice per plane)
        uint32_t col = (uint32_t)((pl & 3) * 96);
        if (MODE == 8 || MODE == 9) {
          col = wring * 32u;
          wring = wring == 12u ? 0u : wring + 1u;
        }
        const uint64_t astage = a0 + (uint64_t)((ROWB == 64 ? (pl & 3) * 12288 : (pl & 1) * 24576) >> 4);
        if (MODE == 3 || MODE == 4) mbar_wait(smem_u32(&bars[3]), 0);
        if (MODE == 4) mbar_wait(smem_u32(&bars[4]), 0);
        if (MODE == 5) {
          if (!tok0) mbar_wait(smem_u32(&bars[3]), 0);
          if (!tok1) mbar_wait(smem_u32(&bars[4]), 0);
          tok0 = mbar_test_wait(smem_u32(&bars[3]), 0);
          tok1 = mbar_test_wait(smem_u32(&bars[4]), 0);
        }
#pragma unroll
        for (int t = 0; t < 9; ++t) {
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k) {
            const uint64_t aoff = (uint64_t)((((t / 3) * 10 + (t % 3)) * ROWB + k * 32) >> 4);
            const uint64_t boff = (uint64_t)((t * 96 * ROWB + k * 32) >> 4);
            if ((MODE == 6 || MODE == 7) && t == 0 && k == 0) {
#pragma unroll
              for (int z = 0; z < 3; ++z)
                umma_bf16(tmem + col + z * 32, astage, b0 + (uint64_t)((z * 32 * ROWB) >> 4),
                          umma_idesc_bf16(128, 32), z == 2 ? 0u : 1u);
            } else {
              umma_bf16(tmem + col, astage + aoff, b0 + boff, umma_idesc_bf16(128, 96), 1u);
            }
          }
        }
        if (MODE >= 1 && MODE != 6 && MODE != 8) umma_commit(smem_u32(&bars[1]));
        if ((MODE >= 2 && MODE != 6 && MODE != 8)) umma_commit(smem_u32(&bars[2]));
      }
      umma_commit(smem_u32(&bars[0]));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bars[0]), 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

template <int ROWB, int MODE>
static void run_loop(int sms, long long* dcyc) {
  const size_t smem_bytes = 196608 + 64 + 1024;
  CK(cudaFuncSetAttribute(loop_kernel<ROWB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  const int planes = 2048;
  for (int fill = 0; fill < 2; ++fill) {
  loop_kernel<ROWB, MODE><<<sms, 128, smem_bytes>>>(planes, dcyc, fill);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("loop kernel error %s\n", cudaGetErrorString(e)); exit(3); }
  std::vector<long long> h(sms);
  CK(cudaMemcpy(h.data(), dcyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (auto v : h) mx = v > mx ? v : mx;
  const int mmas = 9 * (ROWB / 32);
  printf("LOOP rowbytes=%3d mode=%d fill=%s : %.1f cyc/plane (%d MMAs; N=96 floor %d, smem-bound %d)\n", ROWB, MODE,
         fill ? "random" : "zeros", (double)mx / planes, mmas, mmas * 48, mmas * 56);
  }
}


// ---- shared-memory port contention --------------------------------------------------------
// The MMA stream of loop_kernel<64,0> (18 N=96 MMAs per plane, operand reads ~125 B/cycle) runs
// while a second warp writes into an unrelated shared-memory region by
//   bg=1 TMA halo boxes with 64 B rows | bg=2 TMA halo boxes with 128 B rows |
//   bg=3 cp.async 16 B per thread (64 threads) | bg=4 st.shared.v4 (32 threads) | bg=0 nothing.
// Reports MMA cycles/plane and background bytes written per plane.
__global__ void __launch_bounds__(128, 1)
contend_kernel(const __grid_constant__ CUtensorMap tmap64, const __grid_constant__ CUtensorMap tmap128,
               const uint4* gsrc, int planes, int bg, long long* cycles, long long* bgbytes) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + 196608);
  uint32_t* tmem_slot = (uint32_t*)(bars + 12);
  volatile uint32_t* done_flag = (volatile uint32_t*)(bars + 13);
  uint8_t* scratch = smem + 155648;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 196608 / 16; i += blockDim.x) ((uint4*)smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 12; ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_fence_init();
    *done_flag = 0;
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int warp_u = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  constexpr int ROWB = 64;
  if (warp_u == 0) {
    uint64_t a0 = umma_smem_desc<ROWB>(smem_u32(smem));
    a0 &= ~((uint64_t)0x3FFF << 32);
    a0 |= (uint64_t)((10 * ROWB) >> 4) << 32;
    const uint64_t b0 = umma_smem_desc<ROWB>(smem_u32(smem) + 98304);
    long long t0 = clock64();
    if (elect_one()) {
      for (int pl = 0; pl < planes; ++pl) {
        const uint32_t col = (uint32_t)((pl & 3) * 96);
        const uint64_t astage = a0 + (uint64_t)(((pl & 3) * 12288) >> 4);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const uint64_t aoff = (uint64_t)((((t / 3) * 10 + (t % 3)) * ROWB + k * 32) >> 4);
            const uint64_t boff = (uint64_t)((t * 96 * ROWB + k * 32) >> 4);
            umma_bf16(tmem + col, astage + aoff, b0 + boff, umma_idesc_bf16(128, 96), 1u);
          }
        }
      }
      umma_commit(smem_u32(&bars[0]));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bars[0]), 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) {
      cycles[blockIdx.x] = t1 - t0;
      *done_flag = 1;
    }
  } else if (warp_u == 2 || warp_u == 3) {
    long long bytes = 0;
    if (bg == 1 || bg == 2) {
      if (warp_u == 2 && elect_one()) {
        const CUtensorMap* tm = bg == 1 ? &tmap64 : &tmap128;
        const uint32_t box_bytes = bg == 1 ? 180 * 64 : 180 * 128;
        // 4 boxes in flight, round robin over 4 barriers
        uint32_t ph[4] = {0, 0, 0, 0};
        int issued = 0;
        for (int i = 0; i < 4; ++i) {
          mbar_expect_tx(smem_u32(&bars[4 + i]), box_bytes);
          tma_load_5d(smem_u32(scratch), tm, smem_u32(&bars[4 + i]), 0, -1, -1, i % 3, 0);
          ++issued;
        }
        int i = 0;
        while (!*done_flag) {
          mbar_wait(smem_u32(&bars[4 + i]), ph[i]);
          ph[i] ^= 1u;
          bytes += box_bytes;
          mbar_expect_tx(smem_u32(&bars[4 + i]), box_bytes);
          tma_load_5d(smem_u32(scratch), tm, smem_u32(&bars[4 + i]), 0, -1, -1, issued % 3, 0);
          ++issued;
          i = (i + 1) & 3;
        }
        for (int j = 0; j < 4; ++j) {  // drain
          mbar_wait(smem_u32(&bars[4 + i]), ph[i]);
          i = (i + 1) & 3;
        }
        bgbytes[blockIdx.x] = bytes;
      }
    } else if (bg == 3) {
      const int tid = threadIdx.x - 64;  // 0..63
      const uint32_t dst = smem_u32(scratch) + tid * 16;
      int it = 0;
      while (!*done_flag) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)(u * 1024)),
                       "l"(gsrc + ((it * 8 + u) & 1023) * 64 + tid)
                       : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        bytes += 8 * 16;
        ++it;
      }
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      atomicAdd((unsigned long long*)&bgbytes[blockIdx.x], (unsigned long long)bytes);
    } else if (bg == 5 || bg == 6) {
      // tcgen05.ld stream from this warp's TMEM lane quarter: back to back (5) or one 16-column
      // load per ~250 cycles (6), the real epilogue's rate
      uint32_t acc = 0;
      const uint32_t taddr = tmem + ((uint32_t)(warp_u * 32) << 16);
      while (!*done_flag) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
              "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
              "=r"(r[14]), "=r"(r[15])
            : "r"(taddr + ((uint32_t)bytes & 0x1F0u))
            : "memory");
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc += r[j];
        bytes += 16 * 4 * 32;
        if (bg == 6) __nanosleep(100);
      }
      if (acc == 0x12345u) printf("x");
      if ((threadIdx.x & 31) == 0) atomicAdd((unsigned long long*)&bgbytes[blockIdx.x], (unsigned long long)bytes);
    } else if (bg == 4) {
      if (warp_u == 2) {
        const uint32_t dst = smem_u32(scratch) + (threadIdx.x & 31) * 16;
        uint32_t v = threadIdx.x;
        while (!*done_flag) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
            asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst + (uint32_t)(u * 512)), "r"(v) : "memory");
          bytes += 8 * 16;
          ++v;
        }
        atomicAdd((unsigned long long*)&bgbytes[blockIdx.x], (unsigned long long)bytes);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

static void run_contend(int sms, long long* dcyc) {
  // small global tensors for the TMA background stream (L2 resident)
  const int D = 3, H = 24, W = 16;
  uint16_t *d64, *d128;
  uint4* gsrc;
  long long* dbytes;
  CK(cudaMalloc(&d64, (size_t)D * H * W * 32 * 2));
  CK(cudaMalloc(&d128, (size_t)D * H * W * 64 * 2));
  CK(cudaMalloc(&gsrc, 1024 * 64 * 16));
  CK(cudaMalloc(&dbytes, sizeof(long long) * sms));
  CK(cudaMemset(d64, 0, (size_t)D * H * W * 32 * 2));
  CK(cudaMemset(d128, 0, (size_t)D * H * W * 64 * 2));
  CK(cudaMemset(gsrc, 0, 1024 * 64 * 16));
  CUtensorMap t64, t128;
  for (int C : {32, 64}) {
    uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)D, 1};
    uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2, (uint64_t)D * H * W * C * 2};
    uint32_t box[5] = {(uint32_t)C, 10, 18, 1, 1};
    Status st = make_tmap_bf16(C == 32 ? &t64 : &t128, C == 32 ? (void*)d64 : (void*)d128, 5, dims, str, box, C * 2);
    if (!st.ok) { printf("tmap contend: %s\n", st.msg.c_str()); exit(2); }
  }
  const size_t smem_bytes = 196608 + 128 + 1024;
  CK(cudaFuncSetAttribute(contend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  const int planes = 2048;
  for (int bg = 0; bg <= 6; ++bg) {
    CK(cudaMemset(dbytes, 0, sizeof(long long) * sms));
    contend_kernel<<<sms, 128, smem_bytes>>>(t64, t128, gsrc, planes, bg, dcyc, dbytes);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("contend kernel error %s\n", cudaGetErrorString(e)); exit(3); }
    std::vector<long long> h(sms), hb(sms);
    CK(cudaMemcpy(h.data(), dcyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hb.data(), dbytes, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    long long mx = 0; double bsum = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    for (auto v : hb) bsum += (double)v;
    const char* names[7] = {"none", "TMA 64B rows", "TMA 128B rows", "cp.async 16B", "st.shared.v4", "tcgen05.ld b2b", "tcgen05.ld paced"};
    const double cyc = (double)mx / planes, bpp = bsum / sms / planes;
    printf("CONTEND bg=%-14s : %.1f cyc/plane (alone 1008), background %.0f B/plane -> %.1f B per stolen cycle\n",
           names[bg], cyc, bpp, cyc > 1009 ? bpp / (cyc - 1008.2) : 0.0);
  }
}

template <int N, int ACCS>
static void run_rate(int sms, long long* dcyc) {
  const size_t smem_bytes = 196608 + 64 + 1024;
  CK(cudaFuncSetAttribute(rate_kernel<N, ACCS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
  const int iters = 8192;
  for (int grid : {1, sms}) {
    rate_kernel<N, ACCS><<<grid, 128, smem_bytes>>>(iters, dcyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("rate kernel error %s\n", cudaGetErrorString(e)); exit(3); }
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), dcyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (auto v : h) mx = v > mx ? v : mx;
    double cyc = (double)mx / iters;
    printf("RATE accs=%d N=%3d grid=%3d : %.2f cyc/MMA (ideal %.1f) -> %.1f%% of tensor peak\n", ACCS, N,
           grid, cyc, N / 2.0, 100.0 * (N / 2.0) / cyc);
  }
}

// ---- host ------------------------------------------------------------------
static inline float bf16_to_f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
static inline uint16_t f_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return (uint16_t)(u >> 16);  // exact for the small values used here
}

int main() {
  const int D = 3, H = 24, W = 16, NW = 256;
  int failures = 0;
  for (int rb : {128, 64, 32}) {
    if (getenv("PROBE_RATE_ONLY")) break;
    const int C = rb / 2;
    std::vector<uint16_t> hx((size_t)D * H * W * C), hw((size_t)27 * NW * C);
    uint32_t s = 12345u + rb;
    auto rnd = [&]() {
      s = s * 1664525u + 1013904223u;
      return (float)((int)((s >> 16) % 9) - 4) * 0.25f;
    };
    for (auto& v : hx) v = f_to_bf16(rnd());
    for (auto& v : hw) v = f_to_bf16(rnd());
    uint16_t *dx, *dw;
    float* dout;
    CK(cudaMalloc(&dx, hx.size() * 2));
    CK(cudaMalloc(&dw, hw.size() * 2));
    CK(cudaMalloc(&dout, 128 * 256 * 4));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));

    CUtensorMap tx_plain, tx_halo, tw;
    {
      uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)D, 1};
      uint64_t str[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2,
                         (uint64_t)D * H * W * C * 2};
      uint32_t box_plain[5] = {(uint32_t)C, 8, 16, 1, 1};
      uint32_t box_halo[5] = {(uint32_t)C, 10, 18, 1, 1};
      Status st = make_tmap_bf16(&tx_plain, dx, 5, dims, str, box_plain, rb);
      if (!st.ok) { printf("tmap plain: %s\n", st.msg.c_str()); return 2; }
      st = make_tmap_bf16(&tx_halo, dx, 5, dims, str, box_halo, rb);
      if (!st.ok) { printf("tmap halo: %s\n", st.msg.c_str()); return 2; }
      uint64_t wd[3] = {(uint64_t)C, (uint64_t)NW, 27};
      uint64_t ws[2] = {(uint64_t)C * 2, (uint64_t)NW * C * 2};
      uint32_t wb[3] = {(uint32_t)C, 256, 1};
      st = make_tmap_bf16(&tw, dw, 3, wd, ws, wb, rb);
      if (!st.ok) { printf("tmap w: %s\n", st.msg.c_str()); return 2; }
    }
    const size_t smem_bytes = 32768 + 32768 + 64 + 1024;
    if (rb == 128) CK(cudaFuncSetAttribute(probe_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    else if (rb == 64) CK(cudaFuncSetAttribute(probe_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    else CK(cudaFuncSetAttribute(probe_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));

    // tile origin (x0,y0) = (0,0) so the halo box starts at (-1,-1): exercises OOB zero fill,
    // and H=24 > 16+1, W=16 > 8+1 so the far halo is real data.
    auto run = [&](int halo, int ky, int kx, int z, int n, int use_bo, int b_row_off, const char* name, int dcol = 0) {
      ProbeArgs a;
      a.halo = halo; a.ky = ky; a.kx = kx; a.z = z; a.tap = (ky * 3 + kx); a.n = n;
      a.use_base_offset = use_bo; a.b_row_off = b_row_off; a.out = dout; a.dcol = dcol;
      CK(cudaMemset(dout, 0xff, 128 * 256 * 4));
      if (rb == 128) probe_kernel<128><<<1, 128, smem_bytes>>>(halo ? tx_halo : tx_plain, tw, a);
      else if (rb == 64) probe_kernel<64><<<1, 128, smem_bytes>>>(halo ? tx_halo : tx_plain, tw, a);
      else probe_kernel<32><<<1, 128, smem_bytes>>>(halo ? tx_halo : tx_plain, tw, a);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("[rb=%d] %-28s ky=%d kx=%d n=%d : LAUNCH ERROR %s\n", rb, name, ky, kx, n, cudaGetErrorString(e));
        exit(3);
      }
      std::vector<float> ho((size_t)128 * n);
      CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      int bad = 0;
      for (int r = 0; r < 128; ++r) {
        const int ty = r / 8, tx = r % 8;
        const int y = ty + ky - 1, x = tx + kx - 1;
        for (int j = 0; j < n; ++j) {
          float ref = 0.f;
          if (y >= 0 && y < H && x >= 0 && x < W)
            for (int c = 0; c < C; ++c)
              ref += bf16_to_f(hx[(((size_t)z * H + y) * W + x) * C + c]) *
                     bf16_to_f(hw[((size_t)a.tap * NW + b_row_off + j) * C + c]);
          double err = fabs((double)ref - (double)ho[(size_t)r * n + j]);
          if (!(err <= 1e-6)) ++bad;
          if (err > maxerr || err != err) maxerr = err;
        }
      }
      printf("[rb=%3d] %-28s ky=%d kx=%d z=%d n=%3d brow=%3d : max_err=%g bad=%d %s\n", rb, name, ky, kx, z, n,
             b_row_off, maxerr, bad, bad ? "MISMATCH" : "ok");
      return bad;
    };
    // 1. plain path, several N
    for (int n : {32, 64, 96, 128, 256}) failures += run(0, 1, 1, 1, n, 0, 0, "plain") ? 1 : 0;
    failures += run(0, 0, 0, 0, 32, 0, 0, "plain/oob") ? 1 : 0;
    failures += run(0, 2, 2, 2, 64, 0, 32, "plain/b_row_off") ? 1 : 0;
    failures += run(0, 1, 1, 1, 96, 0, 0, "plain/dcol=416", 416) ? 1 : 0;
    // (an accumulator window crossing column 512 does not wrap: the launch faults -- measured)
    // 2. halo view, base_offset = 0
    int halo_bad0 = 0, halo_bad1 = 0;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) halo_bad0 += run(1, ky, kx, 1, 96, 0, 0, "halo/base_offset=0") ? 1 : 0;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) halo_bad1 += run(1, ky, kx, 1, 96, 1, 0, "halo/base_offset=addr") ? 1 : 0;
    printf("[rb=%3d] SUMMARY halo base_offset=0: %d/9 taps bad; base_offset=addr: %d/9 taps bad\n", rb, halo_bad0,
           halo_bad1);
    cudaFree(dx); cudaFree(dw); cudaFree(dout);
  }

  // 3. issue rate
  {
    int dev = 0, sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long* dcyc;
    CK(cudaMalloc(&dcyc, sizeof(long long) * sms));
    run_rate_halo<96, 128, false>(sms, dcyc); run_rate_halo<96, 128, true>(sms, dcyc);
    run_rate_halo<96, 64, false>(sms, dcyc);  run_rate_halo<96, 64, true>(sms, dcyc);
    run_rate_halo<192, 128, false>(sms, dcyc); run_rate_halo<192, 128, true>(sms, dcyc);
    run_contend(sms, dcyc);
    run_loop<64, 0>(sms, dcyc); run_loop<64, 2>(sms, dcyc); run_loop<64, 5>(sms, dcyc); run_loop<64, 7>(sms, dcyc);
    run_loop<64, 8>(sms, dcyc); run_loop<64, 9>(sms, dcyc);
    run_loop<128, 0>(sms, dcyc); run_loop<128, 5>(sms, dcyc);
    if (getenv("PROBE_RATE_ONLY")) { printf("PROBE DONE (rate only)\n"); return 0; }
    run_rate<32, 1>(sms, dcyc);  run_rate<32, 4>(sms, dcyc);
    run_rate<64, 1>(sms, dcyc);  run_rate<64, 4>(sms, dcyc);
    run_rate<96, 1>(sms, dcyc);  run_rate<96, 4>(sms, dcyc);
    run_rate<128, 1>(sms, dcyc); run_rate<128, 4>(sms, dcyc);
    run_rate<192, 1>(sms, dcyc); run_rate<192, 2>(sms, dcyc);
    run_rate<256, 1>(sms, dcyc); run_rate<256, 2>(sms, dcyc);
    cudaFree(dcyc);
  }
  printf("PROBE DONE failures(plain)=%d\n", failures);
  return 0;
}
