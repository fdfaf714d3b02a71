// Host launcher for the tcgen05 implicit-GEMM conv (kernel in conv_umma.cuh).
#include "conv_umma.cuh"

#include <stdlib.h>

#include "kernels.h"
#include "tmap.h"

namespace exa {

namespace {

int largest_pow2_divisor_leq(int value, int cap) {
  int t = 1;
  while (t * 2 <= cap && value % (t * 2) == 0) t *= 2;
  return t;
}

template <int N, int KC, int EPI, int MT = 1>
Status launch_instance(const CUtensorMap& tx, const CUtensorMap& tw, const ConvArgs& a, int grid,
                       cudaStream_t s) {
  using S = ConvSmem<N, KC, MT>;
  static bool configured = false;
  if (!configured) {
    EXA_CUDA(cudaFuncSetAttribute(conv3x3_umma_kernel<N, KC, EPI, MT>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    configured = true;
  }
  conv3x3_umma_kernel<N, KC, EPI, MT><<<grid, 256, S::TOTAL, s>>>(tx, tw, a);
  EXA_CUDA(cudaGetLastError());
  return Status::OK();
}

}  // namespace

Status launch_conv_umma(const Act& in, const Act& out, const __nv_bfloat16* w_packed,
                        const float* bias, const HeadParams* head, int num_sms, cudaStream_t s) {
  EXA_CHECK(!in.fp32, "conv_umma expects bf16 input");
  const int Cin = in.C, Cout = head ? 32 : out.C;
  EXA_CHECK(in.B == out.B || head, "conv_umma: batch mismatch");
  EXA_CHECK(Cin % 32 == 0 && Cout % 32 == 0, "conv_umma: channels must be multiples of 32");
  const int KC = (Cin % 64 == 0) ? 64 : 32;

  int N = Cout;
  if (N > 256) N = 256;
  // two M tiles per stage share one weight tile when the accumulators fit TMEM twice
  // (2 sets x 2 tiles x N <= 512) -- halves the weight traffic per MMA
  static const bool no_mt2 = getenv("EXA_NO_MT2") != nullptr;
  const int mt_try = (!head && !no_mt2 && KC == 64 && N <= 128 && in.voxels() >= 256) ? 2 : 1;

  ConvArgs a{};
  a.B = in.B; a.D = in.D; a.H = in.H; a.W = in.W;
  a.Cin = Cin; a.Cout = Cout;
  int MT = mt_try;
  for (;;) {
    const int rows = 128 * MT;
    a.tw = largest_pow2_divisor_leq(in.W, 32);
    a.th = largest_pow2_divisor_leq(in.H, rows / a.tw);
    a.td = largest_pow2_divisor_leq(in.D, rows / (a.tw * a.th));
    a.tb = rows / (a.tw * a.th * a.td);
    // fall back to single tiles when the box does not fit or the doubled tiles would leave the
    // persistent grid with fewer than ~3 rounds (tile quantisation costs more than B traffic)
    const long long tiles2 = (long long)(in.W / a.tw) * (in.H / a.th) * (in.D / a.td) *
                             ceil_div(in.B, a.tb) * (Cout / N);
    if (MT == 1 || (a.tb <= 256 && tiles2 >= 3LL * num_sms)) break;
    MT = 1;
  }
  a.ntx = in.W / a.tw; a.nty = in.H / a.th; a.ntz = in.D / a.td;
  a.ntb = ceil_div(in.B, a.tb);
  const int m_tiles = a.ntx * a.nty * a.ntz * a.ntb;
  if (N == 256 && m_tiles < num_sms) N = 128;
  EXA_CHECK(Cout % N == 0, "conv_umma: Cout must be a multiple of the N tile");
  a.n_tiles_n = Cout / N;
  a.num_tiles = m_tiles * a.n_tiles_n;
  a.bias = bias;
  if (head) {
    EXA_CHECK(Cout == 32 && N == 32, "fused head needs Cout == 32");
    a.head_w = head->w; a.head_b = head->b; a.head_out = head->out;
    a.head_c = head->C; a.trim = head->trim; a.apply_sigmoid = head->apply_sigmoid;
  } else {
    EXA_CHECK(out.C == Cout && !out.fp32, "conv_umma: output must be bf16 with Cout channels");
    EXA_CHECK(out.D == in.D && out.H == in.H && out.W == in.W, "conv_umma: spatial mismatch");
    EXA_CHECK((out.cstride % 8) == 0 && (out.coff % 8) == 0, "conv_umma: output alignment");
    a.out = (__nv_bfloat16*)out.ptr; a.out_cstride = out.cstride; a.out_coff = out.coff;
  }

  // input activation: 5-D (C, W, H, D, B), box = (KC, tw, th, td, tb), zero OOB fill
  CUtensorMap tx, tw_;
  {
    const uint64_t cs = (uint64_t)in.cstride * 2;
    uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)in.W, (uint64_t)in.H, (uint64_t)in.D,
                        (uint64_t)in.B};
    uint64_t strides[4] = {cs, cs * in.W, cs * in.W * in.H, cs * in.W * in.H * in.D};
    uint32_t box[5] = {(uint32_t)KC, (uint32_t)a.tw, (uint32_t)a.th, (uint32_t)a.td,
                       (uint32_t)a.tb};
    void* base = (void*)((__nv_bfloat16*)in.ptr + in.coff);
    EXA_TRY(make_tmap_bf16(&tx, base, 5, dims, strides, box, KC * 2));
  }
  // weights: 3-D (Cin, Cout, 27), box = (KC, N, 1)
  {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 27};
    uint64_t strides[2] = {(uint64_t)Cin * 2, (uint64_t)Cin * Cout * 2};
    uint32_t box[3] = {(uint32_t)KC, (uint32_t)N, 1};
    EXA_TRY(make_tmap_bf16(&tw_, (void*)w_packed, 3, dims, strides, box, KC * 2));
  }
  const int grid = a.num_tiles < num_sms ? a.num_tiles : num_sms;

  if (MT == 2) {
    if (N == 128) return launch_instance<128, 64, EPI_STORE, 2>(tx, tw_, a, grid, s);
    if (N == 64) return launch_instance<64, 64, EPI_STORE, 2>(tx, tw_, a, grid, s);
    if (N == 32) return launch_instance<32, 64, EPI_STORE, 2>(tx, tw_, a, grid, s);
    return Status::Err("conv_umma: unsupported N for two-tile stages");
  }
#define EXA_CONV_CASE(NN, KK)                                                    \
  if (N == NN && KC == KK) {                                                     \
    if (head) return launch_instance<NN, KK, EPI_HEAD>(tx, tw_, a, grid, s);     \
    return launch_instance<NN, KK, EPI_STORE>(tx, tw_, a, grid, s);              \
  }
#define EXA_CONV_CASE_STORE(NN, KK) \
  if (N == NN && KC == KK) return launch_instance<NN, KK, EPI_STORE>(tx, tw_, a, grid, s);

  EXA_CONV_CASE(32, 32)
  EXA_CONV_CASE(32, 64)
  EXA_CHECK(!head, "fused head only for N == 32");
  EXA_CONV_CASE_STORE(64, 32)
  EXA_CONV_CASE_STORE(64, 64)
  EXA_CONV_CASE_STORE(128, 32)
  EXA_CONV_CASE_STORE(128, 64)
  EXA_CONV_CASE_STORE(256, 32)
  EXA_CONV_CASE_STORE(256, 64)
#undef EXA_CONV_CASE
#undef EXA_CONV_CASE_STORE
  return Status::Err("conv_umma: unsupported (N, KC) = (" + std::to_string(N) + ", " +
                     std::to_string(KC) + ")");
}

}  // namespace exa
