// Host-side launcher of the tcgen05 GEMM (dz_tc.cuh) + a C-ABI self-test entry point.
#include "dz_tc.cuh"
#include "dz_internal.cuh"

namespace dz {

template <int BNJ, int STAGES, bool AK, bool BK, bool AX>
static int launch_tc_t(const char* tag, const TcBatch& tb, void* stream) {
  using L = tc::SmemLayout<BNJ, STAGES>;
  static bool configured = false;
  if (!configured) {
    DZ_CUDA_OK(cudaFuncSetAttribute(tc::tc_gemm_kernel<BNJ, STAGES, AK, BK, AX>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    configured = true;
  }
  int max_i = 0, max_j = 0, max_s = 1;
  for (int q = 0; q < tb.n; ++q) {
    max_i = tb.p[q].MI > max_i ? tb.p[q].MI : max_i;
    max_j = tb.p[q].NJ > max_j ? tb.p[q].NJ : max_j;
    max_s = tb.p[q].splits > max_s ? tb.p[q].splits : max_s;
  }
  dim3 grid((unsigned)ceil_div(max_j, BNJ), (unsigned)(ceil_div(max_i, 128) * max_s), tb.n);
  DZ_LAUNCH_NAMED(tag, (tc::tc_gemm_kernel<BNJ, STAGES, AK, BK, AX>), grid, tc::kThreads, L::kTotal, stream, tb);
  return DZ_OK;
}

template <int BNJ, int STAGES>
static int launch_tc_o(const char* tag, const TcBatch& tb, bool ak, bool bk, bool ax, void* stream) {
  if (ax) {   // exact A operand (uint8 observations): only the two shapes conv1 needs
    if (BNJ == 32 && ak && !bk) return launch_tc_t<32, 4, true, false, true>(tag, tb, stream);    // conv1 forward
    if (BNJ == 32 && !ak && !bk) return launch_tc_t<32, 4, false, false, true>(tag, tb, stream);  // conv1 weight gradient
    return fail(DZ_EINVAL, "exact-A tcgen05 path is instantiated for conv1 only");
  }
  if (ak && bk) return launch_tc_t<BNJ, STAGES, true, true, false>(tag, tb, stream);
  if (ak && !bk) return launch_tc_t<BNJ, STAGES, true, false, false>(tag, tb, stream);
  if (!ak && bk) return launch_tc_t<BNJ, STAGES, false, true, false>(tag, tb, stream);
  return launch_tc_t<BNJ, STAGES, false, false, false>(tag, tb, stream);
}

int launch_tc(const char* tag, const TcBatch& tb_in, int bnj, void* stream) {
  if (tb_in.n <= 0 || tb_in.n > kTcMaxProblems) return fail(DZ_EINVAL, "tc batch size");
  TcBatch tb = tb_in;
  const bool ak = tb.p[0].A.red_is_b != 0, bk = tb.p[0].B.red_is_b != 0, ax = tb.p[0].A.exact != 0;
  for (int q = 0; q < tb.n; ++q) {
    if ((tb.p[q].A.red_is_b != 0) != ak || (tb.p[q].B.red_is_b != 0) != bk || (tb.p[q].A.exact != 0) != ax)
      return fail(DZ_EINVAL, "tc batch mixes operand orientations");
    if (tb.p[q].A.ones_value == 0.f) tb.p[q].A.ones_value = 1.f;
    if (tb.p[q].B.ones_value == 0.f) tb.p[q].B.ones_value = 1.f;
    tc_finalize(tb.p[q].A);
    tc_finalize(tb.p[q].B);
  }
  switch (bnj) {
    case 32: return launch_tc_o<32, 4>(tag, tb, ak, bk, ax, stream);
    case 64: return launch_tc_o<64, 3>(tag, tb, ak, bk, ax, stream);
    case 128: return launch_tc_o<128, 3>(tag, tb, ak, bk, ax, stream);   // 2*128 = 256 TMEM columns
    default: return fail(DZ_EINVAL, "tc tile N must be 32, 64 or 128");
  }
}

}  // namespace dz

using namespace dz;

namespace {
__global__ void u8_to_unit_table_kernel(float* out) {
  dz::pdl_enter(); out[threadIdx.x] = u8_to_unit(threadIdx.x); }
}  // namespace

extern "C" int dz_test_u8_to_unit(float* d_out256, void* stream) {
  DZ_LAUNCH(u8_to_unit_table_kernel, 1, 256, 0, stream, d_out256);
  return DZ_OK;
}

static int g_tc_variant = 0;
extern "C" int dz_test_tc_set_variant(int32_t v) { g_tc_variant = v; return DZ_OK; }

extern "C" int dz_test_tc_gemm(const float* d_A, int32_t a_na, int32_t a_nb, int32_t a_ld, int32_t a_red_is_b,
                               const float* d_B, int32_t b_na, int32_t b_nb, int32_t b_ld, int32_t b_red_is_b,
                               const float* d_scale_a, const float* d_scale_b, int32_t a_ones_row, float* d_C, int32_t MI,
                               int32_t NJ, int32_t R, int64_t sc_i, int64_t sc_j, int32_t splits, int64_t split_stride,
                               int32_t tile_n, void* stream) {
  TcBatch tb;
  memset(&tb, 0, sizeof(tb));
  tb.n = 1;
  tb.variant = g_tc_variant;
  TcProblem& p = tb.p[0];
  p.A.ptr = d_A; p.A.a_mode = A_PLAIN; p.A.na = a_na; p.A.nb = a_nb; p.A.ld = a_ld; p.A.red_is_b = a_red_is_b;
  p.A.scale_r = d_scale_a; p.A.ones_row = a_ones_row;
  p.B.ptr = d_B; p.B.a_mode = A_PLAIN; p.B.na = b_na; p.B.nb = b_nb; p.B.ld = b_ld; p.B.red_is_b = b_red_is_b;
  p.B.scale_r = d_scale_b; p.B.ones_row = -1;
  p.redirect_row = -1;
  p.MI = MI; p.NJ = NJ; p.R = R; p.C = d_C; p.sc_i = sc_i; p.sc_j = sc_j; p.splits = splits; p.split_stride = split_stride;
  return launch_tc("tc_selftest", tb, tile_n, stream);
}
