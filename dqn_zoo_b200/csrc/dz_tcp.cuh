// Packed-operand tcgen05 GEMM for the LARGE contractions of the learner (IQN's 3136->512 layer runs at
// M = batch * tau_samples = 2048 rows per network apply; networks.py:264-292, iqn/agent.py:178-214).
//
//     D[i,j] = sum_r A(i,r) * B(j,r)        fp32 in, fp32-grade out (error-compensated 3xTF32, see dz_tc.cuh)
//
// Unlike dz_tc.cuh (register-path loaders, built for the small implicit-GEMM layers) the hi/lo TF32 split is
// taken OUT of the GEMM: a bandwidth-bound pack kernel writes each operand once as two "tile images"
// (hi and lo parts) that are already in the canonical no-swizzle K-major shared-memory layout of the UMMA
// descriptors, so that the GEMM's producer is ONE thread issuing cp.async.bulk copies (UBLKCP) that complete
// on an mbarrier, the MMA issuer is one thread, and the remaining warps only drain TMEM.
//
// Image layout of an operand with `rows_pad` rows (multiple of the tile height) and `red_pad` reduction
// elements (multiple of 16), RG = rows_pad / 8:
//     float index of element (row, r) = (((r / 16) * RG + row / 8) * 4 + (r % 16) / 4) * 32 + (row % 8) * 4 + r % 4
// i.e. per 16-deep k-block all rows are contiguous, 8-row x 16-byte core matrices, LBO = 128 B (next core
// matrix along the reduction), SBO = 512 B (next 8 rows).  A (TR rows x 16) tile is TR * 64 contiguous bytes.
//
// Accuracy: the tensor core adds into its fp32 accumulator with round-towards-zero (measured: about 2e-8
// relative per accumulation), so an accumulation run is limited to kRunKB k-blocks (128 reduction elements,
// 48 MMA accumulations); the epilogue warps drain each finished run from TMEM and add it into registers with
// ordinary round-to-nearest fp32 adds while the MMA warp already fills the other TMEM buffer.
#pragma once
#include "dz_tc.cuh"
#include "dz_internal.cuh"

namespace dz {

namespace tcp {

using namespace tc;

constexpr int kEpiWarps = 8;
constexpr int kThreadsP = (2 + kEpiWarps) * 32;

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- pack: fp32 matrix -> hi/lo tile images -----------------------------------------------------------------
// One block = one 64-row x 64-deep tile, staged through shared memory so that both the source reads (either
// orientation) and the image writes (512-byte runs) are coalesced.
__global__ void __launch_bounds__(256) tc_pack_kernel(const __grid_constant__ PackBatch pb) {
  dz::pdl_enter();
  __shared__ float tile[64][65];
  int j = 0;
  while (j + 1 < pb.n && (int)blockIdx.x >= pb.job[j + 1].block0) ++j;
  const PackJob& J = pb.job[j];
  const int t = (int)blockIdx.x - J.block0;
  const int row0 = (t % J.tiles_r) * 64, red0 = (t / J.tiles_r) * 64;
  const int tid = threadIdx.x;
  const bool vec = (J.ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(J.src) & 15) == 0);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int idx = tid + q * 256;
    const int a = idx >> 4, b4 = (idx & 15) * 4;         // a: index along the strided source dim, b4: along the contiguous one
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (J.red_contig) {
      const int row = row0 + a, red = red0 + b4;
      if (row < J.rows && red < J.red) {
        const float* s = J.src + (long long)row * J.ld + red;
        if (vec && red + 3 < J.red) { float4 x = *reinterpret_cast<const float4*>(s); v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; }
        else { for (int e = 0; e < 4; ++e) if (red + e < J.red) v[e] = s[e]; }
      }
      if (row == J.ones_row) for (int e = 0; e < 4; ++e) v[e] = red + e < J.red ? 1.f : 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) tile[a][b4 + e] = v[e];
    } else {
      const int red = red0 + a, row = row0 + b4;
      if (red < J.red && row < J.rows) {
        const float* s = J.src + (long long)red * J.ld + row;
        if (vec && row + 3 < J.rows) { float4 x = *reinterpret_cast<const float4*>(s); v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; }
        else { for (int e = 0; e < 4; ++e) if (row + e < J.rows) v[e] = s[e]; }
      }
      if (red < J.red) for (int e = 0; e < 4; ++e) if (row + e == J.ones_row) v[e] = 1.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) tile[b4 + e][a] = v[e];
    }
  }
  __syncthreads();
  const int RG = J.rows_pad >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int id = tid + q * 256;
    const int r = id & 7, c = (id >> 3) & 3, rg = (id >> 5) & 7, kbl = id >> 8;
    const int row = row0 + rg * 8 + r, red = red0 + kbl * 16 + c * 4;
    if (row < J.rows_pad && red < J.red_pad) {
      const float* s = &tile[rg * 8 + r][kbl * 16 + c * 4];
      float4 x = make_float4(s[0], s[1], s[2], s[3]);
      float4 h, l;
      h.x = rn_tf32(x.x); l.x = rn_tf32(x.x - h.x);
      h.y = rn_tf32(x.y); l.y = rn_tf32(x.y - h.y);
      h.z = rn_tf32(x.z); l.z = rn_tf32(x.z - h.z);
      h.w = rn_tf32(x.w); l.w = rn_tf32(x.w - h.w);
      const long long off = ((((long long)(red >> 4) * RG + (row >> 3)) * 4 + c) << 5) + r * 4;
      *reinterpret_cast<float4*>(J.hi + off) = h;
      *reinterpret_cast<float4*>(J.lo + off) = l;
    }
  }
}

// ---- GEMM -----------------------------------------------------------------------------------------------------
template <int BNJ, int EPI>
struct PkSmem {
  static constexpr int kA = 128 * kPkKB * 4, kB = BNJ * kPkKB * 4;     // bytes of one part (hi or lo) of one stage
  static constexpr int kStage = 2 * kA + 2 * kB;
  static constexpr int kStages = EPI ? 2 : 4;      // EPI 1 (short reduction, epilogue-bound): 2 CTAs per SM
  static constexpr int kBars = 1024;
  static constexpr int kTotal = kBars + kStages * kStage;
};

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// grid = (tiles_j, tiles_i * splits, problems); dynamic smem = PkSmem<BNJ, EPI>::kTotal; 320 threads:
//   warp 0: producer (lane 0 issues the bulk copies)      warp 1: TMEM allocator + MMA issuer
//   warps 2..9: epilogue (TMEM lane quarter = warp % 4, column half = (warp - 2) / 4)
// EPI 0: plain output (split partials, or + bias / ReLU), accumulation runs drained into registers.
// EPI 1: IQN embedding epilogue (networks.py:279-284) for a SHORT reduction (one accumulation run, splits == 1):
//        v = relu(acc + bias[j]) -> e0 (fp32, optional);  h = v * mul[i / mul_div][j] -> written straight as the
//        hi/lo tile images of the NEXT GEMMs' operands (rows i, and optionally the transposed rows j).
template <int BNJ, int EPI>
__global__ void __launch_bounds__(kThreadsP, EPI ? 2 : 1) tc_pgemm_kernel(const __grid_constant__ PkBatch batch) {
  dz::pdl_enter();
  extern __shared__ __align__(128) uint8_t smem[];
  using L = PkSmem<BNJ, EPI>;
  constexpr int ST = L::kStages;
  const PkProblem& p = batch.p[blockIdx.z];
  const int tiles_i = (p.MI + 127) / 128;
  const int tile_i = blockIdx.y % tiles_i, split = blockIdx.y / tiles_i;
  const int i0 = tile_i * 128, j0 = blockIdx.x * BNJ;
  if ((int)blockIdx.y >= tiles_i * p.splits || j0 >= p.NJ) return;
  const int per = (p.nkb + p.splits - 1) / p.splits;
  const int kb0 = split * per;
  const int nkb = max(min(p.nkb, kb0 + per) - kb0, 0);
  const int kRunKB = EPI ? (nkb > 0 ? nkb : 1) : batch.run_kb;   // k-blocks per accumulation run
  const int nruns = (nkb + kRunKB - 1) / kRunKB;

  uint64_t* full = reinterpret_cast<uint64_t*>(smem);      // [ST]  bulk copies landed       (tx-count barrier)
  uint64_t* empty = full + ST;                              // [ST]  MMAs consumed the stage
  uint64_t* acc_full = empty + ST;                          // [2]   accumulation run complete in TMEM buffer b
  uint64_t* acc_empty = acc_full + 2;                       // [2]   epilogue drained TMEM buffer b
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint8_t* stage_base = smem + L::kBars;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kTmemCols = EPI ? BNJ : 2 * BNJ;            // EPI 0: two accumulator buffers (512 columns)

  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < ST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
      for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], kEpiWarps); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------------------------------------------------------- producer
    if (elect_one()) {
      const float* a_hi = p.A.hi; const float* a_lo = p.A.lo; const float* b_hi = p.B.hi; const float* b_lo = p.B.lo;
      const long long a_rg = p.A.rg_total, b_rg = p.B.rg_total;
      for (int it = 0; it < nkb; ++it) {
        const int s = it % ST;
        const uint32_t ph = (uint32_t)(it / ST) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        mbar_expect_tx(&full[s], (uint32_t)L::kStage);
        const long long kb = kb0 + it;
        const long long a_off = (kb * a_rg + (i0 >> 3)) * 128;       // floats: 4 chunks x 32 floats per row group
        const long long b_off = (kb * b_rg + (j0 >> 3)) * 128;
        const uint32_t st = smem_u32(stage_base + (size_t)s * L::kStage);
        bulk_g2s(st, a_hi + a_off, L::kA, &full[s]);
        bulk_g2s(st + L::kA, a_lo + a_off, L::kA, &full[s]);
        bulk_g2s(st + 2 * L::kA, b_hi + b_off, L::kB, &full[s]);
        bulk_g2s(st + 2 * L::kA + L::kB, b_lo + b_off, L::kB, &full[s]);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc = make_idesc(128, BNJ, 0, 0);
    for (int it = 0; it < nkb; ++it) {
      const int s = it % ST;
      const uint32_t ph = (uint32_t)(it / ST) & 1u;
      const int run = it / kRunKB, in_run = it - run * kRunKB;
      const int buf = run & 1;
      if (in_run == 0) {
        mbar_wait(&acc_empty[buf], (((uint32_t)run >> 1) & 1u) ^ 1u);   // buffer drained by the epilogue (free on first use)
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      mbar_wait(&full[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {   // elect.sync: ptxas emits the MMAs back to back (no ELECT / BRA.U.ANY loop around each)
        const uint32_t st = smem_u32(stage_base + (size_t)s * L::kStage);
        const uint32_t a_hi = st, a_lo = st + L::kA, b_hi = st + 2 * L::kA, b_lo = st + 2 * L::kA + L::kB;
        const uint32_t d = tmem_base + (uint32_t)(buf * BNJ);
#pragma unroll
        for (int k = 0; k < kPkKB / 8; ++k) {
          const uint64_t dah = make_desc(a_hi + k * 256, 128, 512), dal = make_desc(a_lo + k * 256, 128, 512);
          const uint64_t dbh = make_desc(b_hi + k * 256, 128, 512), dbl = make_desc(b_lo + k * 256, 128, 512);
          mma_tf32(d, dal, dbh, idesc, (in_run > 0 || k > 0) ? 1u : 0u);   // small cross terms first
          mma_tf32(d, dah, dbl, idesc, 1u);
          mma_tf32(d, dah, dbh, idesc, 1u);
        }
        mma_commit(&empty[s]);
        if (in_run == kRunKB - 1 || it == nkb - 1) mma_commit(&acc_full[buf]);
      }
      __syncwarp();
    }
  } else {
    // ---------------------------------------------------------------- epilogue: drain runs, then write
    const int ew = warp - 2;
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, +32) are the ones this warp may read
    const int half = ew >> 2;
    constexpr int kCols = BNJ / 2;                // columns per warp
    if constexpr (EPI == 1) {
      const int i = i0 + quarter * 32 + lane;
      const bool row_ok = i < p.MI;
      const int NJ = p.NJ;
      const float* bias = p.bias_j;
      const float* mulrow = p.mul + (long long)((row_ok ? i : 0) / p.mul_div) * p.mul_ld;
      float* e0 = p.e0 ? p.e0 + (long long)i * p.e0_ld : nullptr;
      float* img_hi = p.img_hi; float* img_lo = p.img_lo; float* imgT_hi = p.imgT_hi; float* imgT_lo = p.imgT_lo;
      const long long rg1 = p.img_rg, rgT = p.imgT_rg;
      if (nkb > 0) {
        mbar_wait(&acc_full[0], 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * kCols);
#pragma unroll 1
      for (int c0 = 0; c0 < kCols; c0 += 32) {
        uint32_t r[32];
        if (nkb > 0) {
          tmem_ld32(taddr + (uint32_t)c0, r);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int t = 0; t < 32; ++t) r[t] = 0u;
        }
#pragma unroll
        for (int t = 0; t < 32; t += 4) {
          const int j = j0 + half * kCols + c0 + t;          // NJ % 4 == 0: a group of four is all valid or all invalid
          const bool ok = row_ok && j < NJ;
          float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok) {
            const float4 b4 = *reinterpret_cast<const float4*>(bias + j);
            float4 v;
            v.x = fmaxf(__uint_as_float(r[t]) + b4.x, 0.f);
            v.y = fmaxf(__uint_as_float(r[t + 1]) + b4.y, 0.f);
            v.z = fmaxf(__uint_as_float(r[t + 2]) + b4.z, 0.f);
            v.w = fmaxf(__uint_as_float(r[t + 3]) + b4.w, 0.f);
            if (e0) *reinterpret_cast<float4*>(e0 + j) = v;
            const float4 m4 = *reinterpret_cast<const float4*>(mulrow + j);
            h = make_float4(v.x * m4.x, v.y * m4.y, v.z * m4.z, v.w * m4.w);
          }
          if (ok) {
            const long long off = (((((long long)(j >> 4) * rg1 + (i >> 3)) << 2) + ((j & 15) >> 2)) << 5) + (i & 7) * 4;
            float4 hh, ll;
            hh.x = rn_tf32(h.x); ll.x = rn_tf32(h.x - hh.x);
            hh.y = rn_tf32(h.y); ll.y = rn_tf32(h.y - hh.y);
            hh.z = rn_tf32(h.z); ll.z = rn_tf32(h.z - hh.z);
            hh.w = rn_tf32(h.w); ll.w = rn_tf32(h.w - hh.w);
            *reinterpret_cast<float4*>(img_hi + off) = hh;
            *reinterpret_cast<float4*>(img_lo + off) = ll;
          }
          if (imgT_hi) {                                       // whole warp: 4x4 transpose inside each lane quad
            const float4 ht = quad_transpose(h, lane);         // lane e: h[m0..m0+3] at column j + e
            const int jj = j + (lane & 3), m0 = i & ~3;
            if (m0 < p.MI && jj < NJ) {
              const long long off = (((((long long)(m0 >> 4) * rgT + (jj >> 3)) << 2) + ((m0 & 15) >> 2)) << 5) + (jj & 7) * 4;
              float4 hh, ll;
              hh.x = rn_tf32(ht.x); ll.x = rn_tf32(ht.x - hh.x);
              hh.y = rn_tf32(ht.y); ll.y = rn_tf32(ht.y - hh.y);
              hh.z = rn_tf32(ht.z); ll.z = rn_tf32(ht.z - hh.z);
              hh.w = rn_tf32(ht.w); ll.w = rn_tf32(ht.w - hh.w);
              *reinterpret_cast<float4*>(imgT_hi + off) = hh;
              *reinterpret_cast<float4*>(imgT_lo + off) = ll;
            }
          }
        }
      }
    } else {
    float sum[kCols];
#pragma unroll
    for (int t = 0; t < kCols; ++t) sum[t] = 0.f;
    for (int run = 0; run < nruns; ++run) {
      const int buf = run & 1;
      mbar_wait(&acc_full[buf], ((uint32_t)run >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * BNJ + half * kCols);
#pragma unroll
      for (int c0 = 0; c0 < kCols; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int t = 0; t < 32; ++t) sum[c0 + t] += __uint_as_float(r[t]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[buf]);
    }
    const int i = i0 + quarter * 32 + lane;
    if (i < p.MI) {
      float* dst = p.C + (long long)split * p.split_stride + (long long)i * p.sc_i;
      const int jb = j0 + half * kCols;
      const bool epi = p.splits == 1;
      const float* bias = epi ? p.bias_j : nullptr;
      const bool relu = epi && p.relu;
      const long long sc_j = p.sc_j;
      const bool v4 = sc_j == 1 && ((reinterpret_cast<uintptr_t>(dst + jb) & 15) == 0) && jb + kCols <= p.NJ;
#pragma unroll
      for (int t = 0; t < kCols; t += 4) {
        float v[4] = {sum[t], sum[t + 1], sum[t + 2], sum[t + 3]};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = jb + t + e;
          if (bias && j < p.NJ) v[e] += bias[j];
          if (relu) v[e] = fmaxf(v[e], 0.f);
        }
        if (v4) {
          *reinterpret_cast<float4*>(dst + jb + t) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = jb + t + e;
            if (j < p.NJ) dst[(long long)j * sc_j] = v[e];
          }
        }
      }
    }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace tcp
}  // namespace dz
