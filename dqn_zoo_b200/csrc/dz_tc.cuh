// tcgen05 / TMEM GEMM for the learner (sm_100a): D[i,j] = sum_r A(i,r) * B(j,r), fp32 operands,
// error-compensated 3xTF32 on the 5th-generation tensor cores so results stay within ~2^-21
// relative of an fp32 FMA evaluation (the parity bar is 1e-5 on losses and gradients):
//
//     x = hi + lo,  hi = x with the 13 low mantissa bits cleared (exactly a TF32 value),  lo = x - hi
//     D += A_hi*B_hi + A_lo*B_hi + A_hi*B_lo            (three tcgen05.mma.kind::tf32 per k-step)
//
// Structure of one CTA (288 threads, one 128 x BNJ output tile, optional split of the reduction):
//   warps 0..7  "loaders": global -> registers (float4, two k-blocks in flight) -> hi/lo split ->
//               st.shared in the canonical no-swizzle UMMA layouts (K-major or MN-major core
//               matrices; implicit im2col, uint8->float/255, per-r scaling and the bias "ones row"
//               are folded into this load) -> fence.proxy.async -> mbarrier arrive.  After the main
//               loop the same warps are the epilogue: tcgen05.ld from TMEM -> global.
//   warp 8      TMEM allocator + single-thread tcgen05.mma issuer; tcgen05.commit releases the smem
//               stage back to the loaders and finally publishes the accumulator.
// No TMA: every operand element has to pass through registers for the hi/lo split anyway, and a
// register-path loader gives implicit GEMM for the convolutions for free.
#pragma once
#include "dz_gemm.cuh"

namespace dz {

struct TcOperand {
  // Source matrix S[a][b] with b the contiguous index (same addressing as GemmProblem's A: plain
  // row-major with leading dimension `ld`, or an implicit im2col view of NHWC float / uint8 rows).
  const void* ptr;
  int a_mode;                 // A_PLAIN / A_CONV_F32 / A_CONV_U8
  int na, nb;                 // extents of a and b
  int ld;                     // plain: elements between consecutive a
  int H, W, Cin, S, OH, OW, seg;
  int red_is_b;               // 1: reduction r = b, tile row = a  (K-major)
                              // 0: reduction r = a, tile row = b  (MN-major)
  const float* scale_r;       // optional scale indexed by the reduction index
  int ones_row;               // tile-row index that reads as 1.0 for every valid r (bias-gradient row), or -1
  FastDiv fd_per, fd_ow, fd_seg;   // conv: OH*OW, OW, seg
  int vec_ok;                 // plain: ld % 4 == 0, nb % 4 == 0 and 16-byte aligned base -> float4 fast path
  // Exact-operand fast path (conv1): the uint8 observations are exactly representable in TF32, so the
  // operand is staged as raw integers 0..255 with NO lo part (2 MMAs per k-step instead of 3) and the 1/255 of
  // networks.py:193 moves onto the other operand (`mul_all`, forward) or the output (`out_scale`, weight grad).
  int exact;                  // 1: values are exact TF32 numbers, skip the hi/lo split
  int u8_raw;                 // A_CONV_U8: deliver (float)byte instead of byte/255
  float mul_all;              // != 0: multiply every element (applied before the split)
  float ones_value;           // value the ones_row reads as (1, or 255 when the output is scaled by 1/255)
};
inline void tc_finalize(TcOperand& o) {
  o.fd_per = make_fastdiv(o.OH * o.OW);
  o.fd_ow = make_fastdiv(o.OW);
  o.fd_seg = make_fastdiv(o.seg);
  o.vec_ok = (o.a_mode != A_PLAIN) || ((o.ld % 4 == 0) && (o.nb % 4 == 0) && ((reinterpret_cast<uintptr_t>(o.ptr) & 15) == 0));
}

struct TcProblem {
  TcOperand A, B;             // MMA "A" (rows i, tile 128) and "B" (rows j, tile BNJ)
  int MI, NJ, R;
  float* C;                   // partial s at C + s*split_stride; element (i,j) at i*sc_i + j*sc_j
  long long sc_i, sc_j, split_stride;
  int splits;
  float out_scale;            // != 0: every output (partial or final) is multiplied by this first
  const float* bias_j;        // optional epilogue (only meaningful with splits == 1): + bias_j[j], then ReLU
  int relu;
  // weight-gradient epilogue (splits == 1): second output C2 = v * s2_i[i] * s2_j[j] (noisy sigma weights) and
  // row `redirect_row` (the bias-gradient row) goes to Cb[j] / Cb2[j] = v * s2_j[j] instead of C.
  float* C2; const float* s2_i; const float* s2_j;
  int redirect_row; float* Cb; float* Cb2;
};

constexpr int kTcMaxProblems = 12;
struct TcBatch {
  TcProblem p[kTcMaxProblems];
  int n;
  int variant;   // debug: bit0 swaps LBO/SBO for MN-major operands
};

namespace tc {

constexpr int kBR = 32;            // reduction elements per k-block (4 MMAs of K = 8)
constexpr int kLoaderWarps = 16;    // 4 per SM sub-partition: enough warps in flight to hide the L2 latency of the loads
constexpr int kEpilogueWarps = 8;   // warps 0..7 drain TMEM (2 per lane quarter)
constexpr int kLoaders = kLoaderWarps * 32;
constexpr int kThreads = kLoaders + 32;
constexpr int kChunkPad = 144;     // 128-byte core matrix + 16 bytes so 8 consecutive chunks hit distinct banks

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One lane of a converged warp.  With elect.sync ptxas knows that exactly one thread runs the guarded region and
// emits the tcgen05.mma / TMA instructions back to back; under `if (lane == 0)` it wraps EVERY such instruction in an
// ELECT / BRA.U.ANY loop over the possibly-active lanes (measured: ~75 cycles per MMA instead of the pipe's 16-32).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n .reg .pred P;\n elect.sync _|P, 0xffffffff;\n selp.u32 %0, 1, 0, P;\n}\n" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// 64-bit shared-memory matrix descriptor, SWIZZLE_NONE ("interleave") canonical layouts.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version for sm_100
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__device__ __forceinline__ uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;    // D format f32
  d |= 2u << 7;    // A format tf32
  d |= 2u << 10;   // B format tf32
  d |= (uint32_t)a_mn_major << 15;
  d |= (uint32_t)b_mn_major << 16;
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Tile geometry in shared memory (bytes).  ROWS = 128 (A) or BNJ (B).
//   K-major : chunk(row, c = r/4)   at (row/8)*SBO + c*LBO + (row%8)*16,  LBO = 144, SBO = 8*144
//   MN-major: chunk(q = row/4, r)   at q*SBO + (r/8)*LBO + (r%8)*16,      SBO = 144, LBO = (ROWS/4)*144
template <int ROWS>
struct TileGeo {
  static constexpr int kBytes = (ROWS / 8) * 8 * kChunkPad;   // identical for both majors: ROWS*32*4 * 144/128
  static constexpr int kKmajLBO = kChunkPad, kKmajSBO = 8 * kChunkPad;
  static constexpr int kMNmajSBO = kChunkPad, kMNmajLBO = (ROWS / 4) * kChunkPad;
};

__device__ __forceinline__ float4 u8x4_to_unit(uchar4 u) {
  return make_float4(u8_to_unit(u.x), u8_to_unit(u.y), u8_to_unit(u.z), u8_to_unit(u.w));
}

// Four consecutive elements along b of S[a][b..b+3] (b % 4 == 0); zeros outside [0,na) x [0,nb).
__device__ __forceinline__ float4 op_load4(const TcOperand& o, int a, int b) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (a >= o.na || b >= o.nb) return v;
  if (o.a_mode == A_PLAIN) {
    const float* src = static_cast<const float*>(o.ptr) + (long long)a * o.ld + b;
    if (b + 3 < o.nb && (o.ld & 3) == 0) return *reinterpret_cast<const float4*>(src);
    v.x = src[0];
    if (b + 1 < o.nb) v.y = src[1];
    if (b + 2 < o.nb) v.z = src[2];
    if (b + 3 < o.nb) v.w = src[3];
    return v;
  }
  int per = o.OH * o.OW;
  int img = a / per, rem = a - img * per;
  int oy = rem / o.OW, ox = rem - oy * o.OW;
  int kh = b / o.seg, kr = b - kh * o.seg;
  long long off = ((long long)(oy * o.S + kh) * o.W + ox * o.S) * o.Cin + kr;
  if (o.a_mode == A_CONV_F32) {
    return *reinterpret_cast<const float4*>(static_cast<const float*>(o.ptr) + (long long)img * o.H * o.W * o.Cin + off);
  }
  uchar4 u = *reinterpret_cast<const uchar4*>(static_cast<const uint8_t* const*>(o.ptr)[img] + off);
  return u8x4_to_unit(u);
}

// hi = rn_tf32(x), lo = rn_tf32(x - hi): round-to-nearest on both keeps the split error zero-mean
// (a truncating split biases every product towards zero).
__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split_store(uint8_t* hi_base, uint8_t* lo_base, int off, float4 v) {
  float4 h, l;
  h.x = rn_tf32(v.x); l.x = rn_tf32(v.x - h.x);
  h.y = rn_tf32(v.y); l.y = rn_tf32(v.y - h.y);
  h.z = rn_tf32(v.z); l.z = rn_tf32(v.z - h.z);
  h.w = rn_tf32(v.w); l.w = rn_tf32(v.w - h.w);
  *reinterpret_cast<float4*>(hi_base + off) = h;
  *reinterpret_cast<float4*>(lo_base + off) = l;
}

// 4x4 transpose across the 4 lanes of an aligned lane quad: on entry lane e holds S[r0+e][b..b+3],
// on exit it holds (S[r0+0..3][b+e]).  Round s: lane R receives from lane (R+s)%4 its component R.
__device__ __forceinline__ float pick(const float4& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }
__device__ __forceinline__ float4 quad_transpose(float4 v, int lane) {
  const int e = lane & 3, base = lane & ~3;
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int src = (e + s) & 3;
    float send = pick(v, (e - s) & 3);                 // sender S sends component (S - s) mod 4 == the receiver's index
    float got = __shfl_sync(0xffffffffu, send, base + src);
    r.x = src == 0 ? got : r.x;                        // came from lane S = (e+s)%4 -> reduction element r0+S
    r.y = src == 1 ? got : r.y;
    r.z = src == 2 ? got : r.z;
    r.w = src == 3 ? got : r.w;
  }
  return r;
}

// Loader-side handling of one operand tile (ROWS x 32): NV float4 per thread.  Both source
// orientations end up in the SAME K-major shared-memory layout (tf32 MN-major descriptors returned
// zeros on this part, and one layout keeps the MMA side trivial): sources contiguous along the
// reduction index (KSRC) are copied chunk for chunk; sources contiguous along the tile-row index are
// transposed 4x4 in registers with warp shuffles first.  Everything that does not depend on the
// k-block (row decomposition, base pointers, shared-memory offsets) is computed once per CTA.
template <int ROWS, bool KSRC>
struct OperandTile {
  static constexpr int kVec = ROWS * kBR / 4;                         // float4 per tile
  static constexpr int NV = (kVec + kLoaders - 1) / kLoaders;         // ROWS=128: 2, 64: 1, 32: 1 (half the threads)
  float4 v[2][NV];          // two k-blocks in flight
  const uint8_t* base[NV];  // KSRC: byte address of S[a][0] (conv: of the pixel's patch origin); else of S[0][b] / unused
  int p0[NV], p1[NV];       // KSRC conv: (kh, kr) of the current k-block; !KSRC: (b or conv k-offset, ones-mask)
  int soff[NV];

  __device__ __forceinline__ void init(const TcOperand& o, int row0, int r_begin, int lt) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      const int idx = lt + q * kLoaders;
      soff[q] = -1; base[q] = nullptr; p0[q] = 0; p1[q] = 0;
      if (idx >= kVec) continue;
      int row, c;
      if (KSRC) { row = idx >> 3; c = idx & 7; }
      else { int e = idx & 3, t = idx >> 2; row = (t % (ROWS / 4)) * 4 + e; c = t / (ROWS / 4); }
      soff[q] = (row >> 3) * TileGeo<ROWS>::kKmajSBO + c * TileGeo<ROWS>::kKmajLBO + (row & 7) * 16;
      if (KSRC) {
        const int a = row0 + row;
        if (a < o.na) {
          if (o.a_mode == A_PLAIN) {
            base[q] = reinterpret_cast<const uint8_t*>(static_cast<const float*>(o.ptr) + (long long)a * o.ld);
          } else {
            int img = fd_div(a, o.fd_per), rem = a - img * o.fd_per.d;
            int oy = fd_div(rem, o.fd_ow), ox = rem - oy * o.fd_ow.d;
            long long off = ((long long)(oy * o.S) * o.W + ox * o.S) * o.Cin;
            if (o.a_mode == A_CONV_F32)
              base[q] = reinterpret_cast<const uint8_t*>(static_cast<const float*>(o.ptr) + (long long)img * o.H * o.W * o.Cin + off);
            else
              base[q] = static_cast<const uint8_t* const*>(o.ptr)[img] + off;
            int b = r_begin + c * 4;
            p0[q] = fd_div(b, o.fd_seg);
            p1[q] = b - p0[q] * o.fd_seg.d;
          }
        }
      } else {
        const int b = row0 + (row & ~3);            // first of the 4 tile rows this float4 covers
        int mask = 0;
        if (o.ones_row >= 0) mask = (b == o.ones_row) | ((b + 1 == o.ones_row) << 1) | ((b + 2 == o.ones_row) << 2) | ((b + 3 == o.ones_row) << 3);
        p1[q] = mask;
        if (o.a_mode == A_PLAIN) {
          p0[q] = b;
        } else {                                     // conv: fixed patch offset of k index b
          int kh = fd_div(b, o.fd_seg), kr = b - kh * o.fd_seg.d;
          p0[q] = b < o.nb ? kh * o.W * o.Cin + kr : -1;
        }
      }
    }
  }

  // Issue the global loads of k-block starting at reduction index r0 into register set `set`.
  template <int set>
  __device__ __forceinline__ void load(const TcOperand& o, int r0, int lt) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      const int idx = lt + q * kLoaders;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx >= kVec) { v[set][q] = x; continue; }
      if (KSRC) {
        const int c = idx & 7;
        const int b = r0 + c * 4;
        if (base[q] && b < o.nb) {
          if (o.a_mode == A_PLAIN) {
            const float* src = reinterpret_cast<const float*>(base[q]) + b;
            if (o.vec_ok) x = *reinterpret_cast<const float4*>(src);
            else { x.x = src[0]; if (b + 1 < o.nb) x.y = src[1]; if (b + 2 < o.nb) x.z = src[2]; if (b + 3 < o.nb) x.w = src[3]; }
          } else {
            const int off = p0[q] * o.W * o.Cin + p1[q];
            if (o.a_mode == A_CONV_F32) x = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base[q]) + off);
            else { uchar4 u = *reinterpret_cast<const uchar4*>(base[q] + off); x = o.u8_raw ? make_float4((float)u.x, (float)u.y, (float)u.z, (float)u.w) : u8x4_to_unit(u); }
          }
          if (o.scale_r) {
            float4 sc = *reinterpret_cast<const float4*>(o.scale_r + b);
            x.x *= sc.x; x.y *= sc.y; x.z *= sc.z; x.w *= sc.w;
          }
        }
        if (o.a_mode != A_PLAIN) {               // advance (kh, kr) by one k-block (seg >= 32)
          p1[q] += kBR;
          if (p1[q] >= o.fd_seg.d) { p1[q] -= o.fd_seg.d; p0[q] += 1; }
        }
      } else {
        const int e = idx & 3, c = (idx >> 2) / (ROWS / 4);
        const int a = r0 + c * 4 + e;              // reduction index (a row of the source)
        if (a < o.na) {
          if (o.a_mode == A_PLAIN) {
            const int b = p0[q];
            if (b < o.nb) {
              const float* src = static_cast<const float*>(o.ptr) + (long long)a * o.ld + b;
              if (o.vec_ok) x = *reinterpret_cast<const float4*>(src);
              else { x.x = src[0]; if (b + 1 < o.nb) x.y = src[1]; if (b + 2 < o.nb) x.z = src[2]; if (b + 3 < o.nb) x.w = src[3]; }
            }
          } else if (p0[q] >= 0) {
            int img = fd_div(a, o.fd_per), rem = a - img * o.fd_per.d;
            int oy = fd_div(rem, o.fd_ow), ox = rem - oy * o.fd_ow.d;
            long long off = ((long long)(oy * o.S) * o.W + ox * o.S) * o.Cin + p0[q];
            if (o.a_mode == A_CONV_F32)
              x = *reinterpret_cast<const float4*>(static_cast<const float*>(o.ptr) + (long long)img * o.H * o.W * o.Cin + off);
            else
            { uchar4 u = *reinterpret_cast<const uchar4*>(static_cast<const uint8_t* const*>(o.ptr)[img] + off); x = o.u8_raw ? make_float4((float)u.x, (float)u.y, (float)u.z, (float)u.w) : u8x4_to_unit(u); }
          }
          if (o.scale_r) { float sc = o.scale_r[a]; x.x *= sc; x.y *= sc; x.z *= sc; x.w *= sc; }
          const int m = p1[q];
          if (m) { const float ov = o.ones_value; if (m & 1) x.x = ov; if (m & 2) x.y = ov; if (m & 4) x.z = ov; if (m & 8) x.w = ov; }
        }
      }
      if (o.mul_all != 0.f) { x.x *= o.mul_all; x.y *= o.mul_all; x.z *= o.mul_all; x.w *= o.mul_all; }
      v[set][q] = x;
    }
  }

  // hi/lo split + store of register set `set` (for !KSRC: 4x4 register transpose first; the shuffles
  // sit here, after the loads have landed, so the load instructions stay back to back).
  template <int set, bool EXACT>
  __device__ __forceinline__ void store(uint8_t* hi, uint8_t* lo, int lt) {
#pragma unroll
    for (int q = 0; q < NV; ++q) {
      if (soff[q] < 0) continue;                 // whole warps at a time (kVec is a multiple of 32)
      float4 x = v[set][q];
      if (!KSRC) x = quad_transpose(x, lt);
      if (EXACT) *reinterpret_cast<float4*>(hi + soff[q]) = x;
      else split_store(hi, lo, soff[q], x);
    }
  }
};

template <int BNJ, int STAGES>
struct SmemLayout {
  static constexpr int kA = TileGeo<128>::kBytes, kB = TileGeo<BNJ>::kBytes;
  static constexpr int kStage = 2 * kA + 2 * kB;      // A_hi, A_lo, B_hi, B_lo
  static constexpr int kBars = 256;
  static constexpr int kTotal = kBars + STAGES * kStage;
};

// grid = (tiles_j, tiles_i * splits, problems); dynamic smem = SmemLayout<BNJ,STAGES>::kTotal.
// A_KSRC / B_KSRC: the operand's source is contiguous along the reduction index (uniform over the batch).
template <int BNJ, int STAGES, bool A_KSRC, bool B_KSRC, bool A_EXACT>
__global__ void __launch_bounds__(kThreads, 1) tc_gemm_kernel(const __grid_constant__ TcBatch batch) {
  dz::pdl_enter();
  extern __shared__ __align__(128) uint8_t smem[];
  using L = SmemLayout<BNJ, STAGES>;
  const TcProblem& p = batch.p[blockIdx.z];
  const int tiles_i = (p.MI + 127) / 128;
  const int tile_i = blockIdx.y % tiles_i, split = blockIdx.y / tiles_i;
  const int i0 = tile_i * 128, j0 = blockIdx.x * BNJ;
  if ((int)blockIdx.y >= tiles_i * p.splits || j0 >= p.NJ) return;
  const int nkb_all = (p.R + kBR - 1) / kBR;
  const int per = (nkb_all + p.splits - 1) / p.splits;
  const int kb0 = split * per, kb1 = min(nkb_all, kb0 + per);
  const int nkb = max(kb1 - kb0, 0);

  uint64_t* full = reinterpret_cast<uint64_t*>(smem);        // [STAGES] loaders -> mma
  uint64_t* empty = full + STAGES;                            // [STAGES] mma -> loaders
  uint64_t* accum = empty + STAGES;                           // [1]      mma -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum + 1);
  uint8_t* stage_base = smem + L::kBars;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Two accumulators: columns [0,BNJ) take the hi*hi products, [BNJ,2BNJ) the two small cross terms.  The
  // tensor core adds into its fp32 accumulator with round-towards-zero, so every accumulation step costs
  // ~half an ulp of the running sum; keeping the (2^-11 times smaller) cross terms out of the main
  // accumulator cuts the number of such steps on it by 3x.
  constexpr int kTmemCols = 2 * BNJ < 32 ? 32 : 2 * BNJ;      // power of two >= 32

  if (warp == kLoaderWarps) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], kLoaderWarps); mbar_init(&empty[s], 1); }
      mbar_init(accum, 1);
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

  if (warp < kLoaderWarps) {
    // ------------------------------------------------------------------ loaders
    const int lt = threadIdx.x;
    // Register copies of the operand descriptors: the mbarrier / fence asm statements below carry "memory"
    // clobbers, and fields read through `p` (kernel-parameter space, runtime-indexed) would be re-fetched
    // after every one of them.
    const TcOperand opA = p.A, opB = p.B;
    OperandTile<128, A_KSRC> ta;
    OperandTile<BNJ, B_KSRC> tb;
    ta.init(opA, i0, kb0 * kBR, lt);
    tb.init(opB, j0, kb0 * kBR, lt);
    if (nkb > 0) { ta.template load<0>(opA, kb0 * kBR, lt); tb.template load<0>(opB, kb0 * kBR, lt); }
    if (nkb > 1) { ta.template load<1>(opA, (kb0 + 1) * kBR, lt); tb.template load<1>(opB, (kb0 + 1) * kBR, lt); }
    for (int it = 0; it < nkb; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(&empty[s], ph ^ 1u);
      uint8_t* st = stage_base + (size_t)s * L::kStage;
      const bool more = it + 2 < nkb;
      const int rn = (kb0 + it + 2) * kBR;
      if ((it & 1) == 0) {
        ta.template store<0, A_EXACT>(st, st + L::kA, lt);
        tb.template store<0, false>(st + 2 * L::kA, st + 2 * L::kA + L::kB, lt);
      } else {
        ta.template store<1, A_EXACT>(st, st + L::kA, lt);
        tb.template store<1, false>(st + 2 * L::kA, st + 2 * L::kA + L::kB, lt);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
      if (more) {
        if ((it & 1) == 0) { ta.template load<0>(opA, rn, lt); tb.template load<0>(opB, rn, lt); }
        else               { ta.template load<1>(opA, rn, lt); tb.template load<1>(opB, rn, lt); }
      }
    }
    // ------------------------------------------------------------------ epilogue: TMEM -> global
    if (warp < kEpilogueWarps) {
    if (nkb > 0) {
      mbar_wait(accum, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const int MI = p.MI, NJ = p.NJ, redirect_row = p.redirect_row, relu = p.relu;
    const long long sc_i = p.sc_i, sc_j = p.sc_j;
    const float out_scale = p.out_scale;
    const float* bias_j = p.bias_j; const float* s2_i = p.s2_i; const float* s2_j = p.s2_j;
    float* C2 = p.C2; float* Cb = p.Cb; float* Cb2 = p.Cb2;
    const int quarter = warp & 3;                 // TMEM lanes [32*quarter, 32*quarter+32)
    const int half = warp >> 2;                   // column halves
    constexpr int kColsPerWarp = BNJ / 2;
    const int i = i0 + quarter * 32 + lane;
    float* dst = p.C + (long long)split * p.split_stride;
#pragma unroll
    for (int c0 = 0; c0 < kColsPerWarp; c0 += 16) {
      const int col = half * kColsPerWarp + c0;
      uint32_t r[16], r2[16];
      if (nkb > 0) {
        uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r2[0]), "=r"(r2[1]), "=r"(r2[2]), "=r"(r2[3]), "=r"(r2[4]), "=r"(r2[5]), "=r"(r2[6]), "=r"(r2[7]), "=r"(r2[8]),
              "=r"(r2[9]), "=r"(r2[10]), "=r"(r2[11]), "=r"(r2[12]), "=r"(r2[13]), "=r"(r2[14]), "=r"(r2[15])
            : "r"(taddr + (uint32_t)BNJ));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int t = 0; t < 16; ++t) r[t] = __float_as_uint(__uint_as_float(r[t]) + __uint_as_float(r2[t]));
      } else {
#pragma unroll
        for (int t = 0; t < 16; ++t) r[t] = 0u;
      }
      if (i < MI) {
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          int j = j0 + col + t;
          if (j < NJ) {
            float v = __uint_as_float(r[t]);
            if (out_scale != 0.f) v *= out_scale;
            if (bias_j) v += bias_j[j];
            if (relu) v = fmaxf(v, 0.f);
            if (redirect_row >= 0 && i == redirect_row) {
              if (Cb) Cb[j] = v;
              if (Cb2) Cb2[j] = v * s2_j[j];
            } else {
              const long long at = (long long)i * sc_i + (long long)j * sc_j;
              if (dst) dst[at] = v;
              if (C2) C2[at] = v * s2_i[i] * s2_j[j];
            }
          }
        }
      }
    }
    }
  } else {
    // ------------------------------------------------------------------ MMA issuer (last warp)
    const uint32_t idesc = make_idesc(128, BNJ, 0, 0);          // both operands K-major in shared memory
    const uint32_t a_lbo = TileGeo<128>::kKmajLBO, a_sbo = TileGeo<128>::kKmajSBO;
    const uint32_t b_lbo = TileGeo<BNJ>::kKmajLBO, b_sbo = TileGeo<BNJ>::kKmajSBO;
    const uint32_t a_step = 2 * TileGeo<128>::kKmajLBO, b_step = 2 * TileGeo<BNJ>::kKmajLBO;   // 8 reduction elements = 2 chunks
    for (int it = 0; it < nkb; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
      mbar_wait(&full[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t st = smem_u32(stage_base + (size_t)s * L::kStage);
        const uint32_t a_hi = st, a_lo = st + L::kA, b_hi = st + 2 * L::kA, b_lo = st + 2 * L::kA + L::kB;
#pragma unroll
        for (int k = 0; k < kBR / 8; ++k) {
          uint64_t dah = make_desc(a_hi + k * a_step, a_lbo, a_sbo);
          uint64_t dal = make_desc(a_lo + k * a_step, a_lbo, a_sbo);
          uint64_t dbh = make_desc(b_hi + k * b_step, b_lbo, b_sbo);
          uint64_t dbl = make_desc(b_lo + k * b_step, b_lbo, b_sbo);
          const uint32_t acc = (it > 0 || k > 0) ? 1u : 0u;
          mma_tf32(tmem_base, dah, dbh, idesc, acc);                 // main accumulator
          if (A_EXACT) {
            mma_tf32(tmem_base + BNJ, dah, dbl, idesc, acc);         // A has no lo part
          } else {
            mma_tf32(tmem_base + BNJ, dal, dbh, idesc, acc);         // cross terms
            mma_tf32(tmem_base + BNJ, dah, dbl, idesc, 1u);
          }
        }
        mma_commit(&empty[s]);                    // arrives when the MMAs above have consumed this stage
        if (it == nkb - 1) mma_commit(accum);     // accumulator complete
      }
      __syncwarp();
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kLoaderWarps) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

}  // namespace tc
}  // namespace dz
