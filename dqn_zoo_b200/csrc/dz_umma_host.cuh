// Host side of the TMA-fed tcgen05 GEMM family (dz_umma.cuh): tensor-map encoding and the per-launch tables
// (CTA descriptors, TMA programs).  A UmPlan is built once per learner and replayed every step.
#pragma once
#include <vector>

#include "dz_umma.cuh"

namespace dz {

// One launch of umma_gemm_kernel: a contiguous range of CTA descriptors sharing NJT / stage geometry.
struct UmLaunch {
  int cta0 = 0, nctas = 0;
  int njt = 64;
  int stages = 4;
  uint32_t stage_bytes = 0;
  bool convert = false;
  int nmaps = 0;
  int map_ids[um::kMaxMapsPerLaunch] = {0};   // plan map index of the launch-local map slot (the ops of this launch use slots)
};

struct UmPlan {
  std::vector<CUtensorMap> maps;
  std::vector<UmProblem> probs;
  std::vector<UmCta> ctas;
  std::vector<UmTmaOp> ops;
  // device copies
  CUtensorMap* d_maps = nullptr;
  UmProblem* d_probs = nullptr;
  UmCta* d_ctas = nullptr;
  UmTmaOp* d_ops = nullptr;

  // 5-D fp32 tensor map with SWIZZLE_128B; dims/strides innermost first (strides in BYTES for dims 1..4; dims beyond
  // `rank` are 1).  mn_major: the tile feeds an MN-major (transposing) descriptor -> 32-byte-atom flavour of the swizzle.
  // Returns the map index or -1 (error string set).
  int add_map(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, bool mn_major = false);
  // Rewrites the `map` field of the ops of launch `l` (ctas [cta0, cta0 + nctas)) from plan indices to launch-local slots.
  int localize_maps(UmLaunch& l);
  int upload();          // (re)allocates and copies all four tables
  void release();
  int launch(const char* tag, const UmLaunch& l, void* stream, long long* d_trace = nullptr) const;   // d_trace: 512 clock stamps of CTA 0 (debug)
  static int configure();   // one-time kernel attributes (outside any stream capture)
};

// Operand format helpers
inline UmOperand um_kmajor(int rows, bool hi_lo, bool convert, const float* scale_r = nullptr) {
  UmOperand o;
  memset(&o, 0, sizeof(o));
  o.part_bytes = (uint32_t)(((rows * 128) + 1023) / 1024 * 1024);
  o.nparts = hi_lo ? 2 : 1; o.convert = convert ? 1 : 0; o.mn_major = 0; o.lbo = 0; o.kstep = 32; o.scale_r = scale_r;
  return o;
}
inline UmOperand um_mnmajor(int mn, int r_rows, bool convert, const float* scale_r = nullptr) {
  UmOperand o;
  memset(&o, 0, sizeof(o));
  o.lbo = (uint32_t)(r_rows * 128);
  o.part_bytes = (uint32_t)((mn / 32) * o.lbo);
  o.nparts = 2; o.convert = convert ? 1 : 0; o.mn_major = 1; o.kstep = 1024; o.scale_r = scale_r;
  return o;
}

}  // namespace dz
