// Host side of the TMA-fed tcgen05 GEMM family + a C-ABI self test (plain GEMMs through every operand path).
#include <cudaTypedefs.h>

#include "dz_umma_host.cuh"

namespace dz {

namespace {

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

}  // namespace

int UmPlan::add_map(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, bool mn_major) {
  auto fn = encode_fn();
  if (!fn) { fail(DZ_ECUDA, "cuTensorMapEncodeTiled entry point not available"); return -1; }
  cuuint64_t gd[5] = {1, 1, 1, 1, 1};
  cuuint64_t gs[4] = {16, 16, 16, 16};
  cuuint32_t bx[5] = {1, 1, 1, 1, 1};
  cuuint32_t es[5] = {1, 1, 1, 1, 1};
  uint64_t last = 16;
  for (int i = 0; i < 5; ++i) {
    if (i < rank) { gd[i] = dims[i]; bx[i] = box[i]; }
    if (i >= 1) {
      if (i < rank) gs[i - 1] = strides_bytes[i - 1];
      else gs[i - 1] = last;                  // size-1 dimension: any legal stride
      last = gs[i - 1] * gd[i];
      if (last % 16) last = (last + 15) / 16 * 16;
    } else {
      last = gd[0] * 4;
      if (last % 16) last = (last + 15) / 16 * 16;
    }
  }
  if (bx[0] * 4 > 128) { fail(DZ_EINVAL, "tensor map: inner box wider than the 128-byte swizzle span"); return -1; }
  CUtensorMap m;
  CUresult rc = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu %llu %llu strides %llu %llu %llu %llu box %u %u %u %u %u", (int)rc,
             (unsigned long long)gd[0], (unsigned long long)gd[1], (unsigned long long)gd[2], (unsigned long long)gd[3], (unsigned long long)gd[4],
             (unsigned long long)gs[0], (unsigned long long)gs[1], (unsigned long long)gs[2], (unsigned long long)gs[3], bx[0], bx[1], bx[2], bx[3], bx[4]);
    g_last_error = buf;
    return -1;
  }
  maps.push_back(m);
  return (int)maps.size() - 1;
}

int UmPlan::localize_maps(UmLaunch& l) {
  l.nmaps = 0;
  for (int ci = l.cta0; ci < l.cta0 + l.nctas; ++ci) {
    const UmCta& c = ctas[ci];
    const size_t n = (size_t)c.nstages * c.ops_per_stage;
    for (size_t oi = c.op0; oi < c.op0 + n; ++oi) {
      const int id = (int)ops[oi].map;
      int slot = -1;
      for (int q = 0; q < l.nmaps; ++q) if (l.map_ids[q] == id) slot = q;
      if (slot < 0) {
        if (l.nmaps >= um::kMaxMapsPerLaunch) return fail(DZ_EINVAL, "too many tensor maps in one launch");
        slot = l.nmaps;
        l.map_ids[l.nmaps++] = id;
      }
      ops[oi].map = (uint32_t)slot;
    }
  }
  return DZ_OK;
}

void UmPlan::release() {
  if (d_maps) cudaFree(d_maps);
  if (d_probs) cudaFree(d_probs);
  if (d_ctas) cudaFree(d_ctas);
  if (d_ops) cudaFree(d_ops);
  d_maps = nullptr; d_probs = nullptr; d_ctas = nullptr; d_ops = nullptr;
}

int UmPlan::upload() {
  release();
  if (maps.empty() || probs.empty() || ctas.empty() || ops.empty()) return fail(DZ_EINVAL, "empty umma plan");
  DZ_CUDA_OK(cudaMalloc(&d_maps, maps.size() * sizeof(CUtensorMap)));
  DZ_CUDA_OK(cudaMalloc(&d_probs, probs.size() * sizeof(UmProblem)));
  DZ_CUDA_OK(cudaMalloc(&d_ctas, ctas.size() * sizeof(UmCta)));
  DZ_CUDA_OK(cudaMalloc(&d_ops, ops.size() * sizeof(UmTmaOp)));
  DZ_CUDA_OK(cudaMemcpy(d_maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
  DZ_CUDA_OK(cudaMemcpy(d_probs, probs.data(), probs.size() * sizeof(UmProblem), cudaMemcpyHostToDevice));
  DZ_CUDA_OK(cudaMemcpy(d_ctas, ctas.data(), ctas.size() * sizeof(UmCta), cudaMemcpyHostToDevice));
  DZ_CUDA_OK(cudaMemcpy(d_ops, ops.data(), ops.size() * sizeof(UmTmaOp), cudaMemcpyHostToDevice));
  return DZ_OK;
}

#define DZ_TRY_CFG(expr) do { int _s = (expr); if (_s != DZ_OK) return _s; } while (0)

int UmPlan::configure() {
  static bool done = false;
  if (done) return DZ_OK;
  DZ_CUDA_OK(cudaFuncSetAttribute(um::umma_gemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  DZ_CUDA_OK(cudaFuncSetAttribute(um::umma_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  done = true;
  return DZ_OK;
}

int UmPlan::launch(const char* tag, const UmLaunch& l, void* stream, long long* d_trace) const {
  if (l.nctas <= 0) return DZ_OK;
  if (!d_ctas) return fail(DZ_EINVAL, "umma plan not uploaded");
  if (l.stages < 1 || l.stages > um::kStagesMax || l.stage_bytes % 1024) return fail(DZ_EINVAL, "umma launch geometry");
  for (int ci = l.cta0; ci < l.cta0 + l.nctas; ++ci) {
    const UmProblem& pr = probs[ctas[ci].prob];
    const uint32_t want = pr.B.mn_major ? (uint32_t)(l.njt / 32) * pr.B.lbo : (uint32_t)l.njt * 128u;
    if (pr.A.nparts != 2 || pr.B.nparts != 2 || pr.B.part_bytes != want)
      return fail(DZ_EINVAL, "umma launch: operands must be hi/lo pairs with the B parts adjacent (one descriptor spans both)");
  }
  for (int ci = l.cta0; ci < l.cta0 + l.nctas; ++ci)
    if (ctas[ci].ops_per_stage > 32) return fail(DZ_EINVAL, "umma launch: more than 32 TMA ops per stage");
  // two MMA-issuer warps need static slot ownership: stage count a multiple of 2 * run_stages (dz_umma.cuh)
  int stages = l.stages;
  {
    uint32_t rs = 1;
    for (int ci = l.cta0; ci < l.cta0 + l.nctas; ++ci) rs = std::max(rs, probs[ctas[ci].prob].run_stages);
    const int rounded = (int)(l.stages / (2 * rs) * (2 * rs));
    if (rounded >= (int)(2 * rs)) stages = rounded;
  }
  const size_t smem = 1024 + um::kCtlBytes + (size_t)stages * l.stage_bytes;
  if (smem > 227 * 1024) return fail(DZ_EINVAL, "umma launch needs too much shared memory");
  if ((size_t)stages * l.stage_bytes < (size_t)128 * l.njt * 4) return fail(DZ_EINVAL, "umma launch: stage buffers smaller than the store-phase staging tile");
  const int v = l.njt == 32 ? 0 : 1;
  if (l.njt != 32 && l.njt != 64) return fail(DZ_EINVAL, "umma launch: NJT must be 32 or 64");
  DZ_TRY_CFG(configure());
  um::UmMaps lm;
  for (int q = 0; q < l.nmaps; ++q) lm.m[q] = maps[l.map_ids[q]];
  for (int q = l.nmaps; q < um::kMaxMapsPerLaunch; ++q) lm.m[q] = maps[l.map_ids[0]];
  if (v == 0)
    DZ_LAUNCH_NAMED(tag, um::umma_gemm_kernel<32>, (unsigned)l.nctas, um::kThreadsG, smem, stream, lm, d_ctas + l.cta0, d_probs, d_ops, l.nmaps, stages,
                    l.stage_bytes, d_trace);
  else
    DZ_LAUNCH_NAMED(tag, um::umma_gemm_kernel<64>, (unsigned)l.nctas, um::kThreadsG, smem, stream, lm, d_ctas + l.cta0, d_probs, d_ops, l.nmaps, stages,
                    l.stage_bytes, d_trace);
  return DZ_OK;
}

namespace {
__global__ void um_split_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long long n) {
  dz::pdl_enter();
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) {
    float h = tc::rn_tf32(x[i]);
    hi[i] = h;
    lo[i] = tc::rn_tf32(x[i] - h);
  }
}
}  // namespace

int um_split(const float* x, float* hi, float* lo, long long n, void* stream) {
  DZ_LAUNCH(um_split_kernel, (unsigned)ceil_div(n, 256), 256, 0, stream, x, hi, lo, n);
  return DZ_OK;
}

}  // namespace dz

using namespace dz;

// Self test: C[MI][NJ] = sum_r A(i,r) B(j,r).
//   a_mn_major = 0: d_A is [MI][R] (K-major source);  1: d_A is [R][MI] (MN-major source).  Same for B with NJ <= 64.
//   convert = 0: operands are split into hi/lo by a helper kernel first (the layout the activations use);
//   convert = 1: raw fp32 tiles are split in shared memory by the converter warps (the layout the weights use),
//                optionally scaled by d_scale_r[r] (applied to A).
//   epi_rows = 1: UM_EPI_ROWS epilogue (+ d_bias[j], relu, tf32 hi/lo outputs in d_hi / d_lo besides d_C).
extern "C" int dz_test_umma_gemm(const float* d_A, int32_t a_mn_major, const float* d_B, int32_t b_mn_major, int32_t MI, int32_t NJ,
                                 int32_t R, int32_t convert, const float* d_scale_r, int32_t run_stages, int32_t epi_rows,
                                 const float* d_bias, int32_t relu, float* d_C, float* d_hi, float* d_lo, void* stream) {
  if (NJ < 4 || NJ > 64 || NJ % 4 || MI % 4 || R % 4) return fail(DZ_EINVAL, "umma self test extents");
  const int njt = NJ <= 32 ? 32 : 64;
  UmPlan plan;
  float *a_hi = nullptr, *a_lo = nullptr, *b_hi = nullptr, *b_lo = nullptr;
  const long long na = (long long)MI * R, nb = (long long)NJ * R;
  if (!convert) {
    DZ_CUDA_OK(cudaMalloc(&a_hi, na * 4)); DZ_CUDA_OK(cudaMalloc(&a_lo, na * 4));
    DZ_CUDA_OK(cudaMalloc(&b_hi, nb * 4)); DZ_CUDA_OK(cudaMalloc(&b_lo, nb * 4));
    int rc = um_split(d_A, a_hi, a_lo, na, stream);
    if (rc == DZ_OK) rc = um_split(d_B, b_hi, b_lo, nb, stream);
    if (rc != DZ_OK) return rc;
  }
  auto make_maps = [&](const float* hi, const float* lo, int mn_major, int rows, int tile_rows, int out[2]) -> int {
    const float* src[2] = {hi, lo};
    for (int part = 0; part < (convert ? 1 : 2); ++part) {
      uint64_t dims[2], strides[1];
      uint32_t box[2];
      if (!mn_major) { dims[0] = (uint64_t)R; dims[1] = (uint64_t)rows; strides[0] = (uint64_t)R * 4; box[0] = 32; box[1] = (uint32_t)tile_rows; }
      else { dims[0] = (uint64_t)rows; dims[1] = (uint64_t)R; strides[0] = (uint64_t)rows * 4; box[0] = 32; box[1] = 32; }
      out[part] = plan.add_map(src[part], 2, dims, strides, box, mn_major != 0);
      if (out[part] < 0) return DZ_EINVAL;
    }
    return DZ_OK;
  };
  int ma[2] = {-1, -1}, mb[2] = {-1, -1};
  int rc = make_maps(convert ? d_A : a_hi, a_lo, a_mn_major, MI, 128, ma);
  if (rc == DZ_OK) rc = make_maps(convert ? d_B : b_hi, b_lo, b_mn_major, NJ, njt, mb);
  if (rc != DZ_OK) return rc;

  UmProblem p;
  memset(&p, 0, sizeof(p));
  p.A = a_mn_major ? um_mnmajor(128, 32, convert != 0, convert ? d_scale_r : nullptr) : um_kmajor(128, true, convert != 0, convert ? d_scale_r : nullptr);
  p.B = b_mn_major ? um_mnmajor(njt, 32, convert != 0) : um_kmajor(njt, true, convert != 0);
  p.ksteps = 4; p.run_stages = run_stages > 0 ? run_stages : 1; p.red_per_stage = 32;
  p.MI = MI; p.NJ = NJ;
  if (epi_rows) {
    p.epi = UM_EPI_ROWS; p.out_f32 = d_C; p.out_hi = d_hi; p.out_lo = d_lo; p.bias = d_bias; p.relu = relu;
    p.pw = 1 << 20; p.rs_outer = 0; p.rs_inner = 1; p.out_ld = NJ;
  } else {
    p.epi = UM_EPI_PARTIAL; p.C = d_C; p.sc_i = NJ; p.sc_j = 1; p.split_stride = 0;
  }
  plan.probs.push_back(p);
  const int nst = (int)ceil_div(R, 32);
  const uint32_t a_bytes = p.A.part_bytes * 2, b_bytes = p.B.part_bytes * 2;
  const int tiles = (int)ceil_div(MI, 128);
  for (int t = 0; t < tiles; ++t) {
    UmCta c;
    memset(&c, 0, sizeof(c));
    c.prob = 0; c.op0 = (uint32_t)plan.ops.size(); c.nstages = (uint32_t)nst; c.r0 = 0; c.i0 = t * 128; c.split = 0;
    c.row_base = t * 128; c.ph_valid = 1; c.pw_valid = std::min(128, MI - t * 128);
    uint32_t tx = 0;
    int nops = 0;
    for (int s = 0; s < nst; ++s) {
      nops = 0; tx = 0;
      auto add = [&](int map, uint32_t off, int c0, int c1, uint32_t bytes) {
        UmTmaOp o;
        memset(&o, 0, sizeof(o));
        o.map = (uint32_t)map; o.smem_off = off; o.c[0] = c0; o.c[1] = c1;
        plan.ops.push_back(o);
        ++nops; tx += bytes;
      };
      for (int part = 0; part < (convert ? 1 : 2); ++part) {
        const uint32_t base = part * p.A.part_bytes;
        if (!a_mn_major) add(ma[part], base, 32 * s, t * 128, 128 * 128);
        else for (int q = 0; q < 4; ++q) add(ma[part], base + q * 4096, t * 128 + 32 * q, 32 * s, 4096);
      }
      for (int part = 0; part < (convert ? 1 : 2); ++part) {
        const uint32_t base = a_bytes + part * p.B.part_bytes;
        if (!b_mn_major) add(mb[part], base, 32 * s, 0, (uint32_t)njt * 128);
        else for (int q = 0; q < njt / 32; ++q) add(mb[part], base + q * 4096, 32 * q, 32 * s, 4096);
      }
    }
    c.ops_per_stage = (uint32_t)nops; c.tx_bytes = tx;
    plan.ctas.push_back(c);
  }
  UmLaunch l;
  l.cta0 = 0; l.nctas = tiles;
  rc = plan.localize_maps(l);
  if (rc == DZ_OK) rc = plan.upload();
  if (rc != DZ_OK) return rc;
  l.njt = njt; l.stage_bytes = a_bytes + b_bytes; l.stages = 4; l.convert = convert != 0;   // 4 x <= 48 KB + control block fits
  rc = plan.launch("umma_selftest", l, stream);
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  plan.release();
  if (a_hi) { cudaFree(a_hi); cudaFree(a_lo); cudaFree(b_hi); cudaFree(b_lo); }
  if (rc != DZ_OK) return rc;
  if (e != cudaSuccess) return fail(DZ_ECUDA, "umma self test: %s", cudaGetErrorString(e));
  return DZ_OK;
}
