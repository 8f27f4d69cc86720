// tcgen05 / TMEM helpers shared by the tensor-core kernels of the learner (sm_100a): mbarrier and elect.sync wrappers,
// the kind::tf32 MMA / commit instructions, instruction- and (no-swizzle) shared-memory descriptors, and the tf32 split
//
//     x = hi + lo,  hi = rn_tf32(x),  lo = rn_tf32(x - hi)      (error-compensated 3xTF32: D += Ah*Bh + Al*Bh + Ah*Bl)
//
// that keeps every contraction within ~2^-21 relative of an fp32 FMA evaluation (the parity bar is 1e-5 on losses and
// gradients).  The kernels themselves: dz_umma.cuh (TMA-fed family: torso + 3136->512 layers at batch 32), dz_umma_net.cu
// (conv1 with the uint8 gather), dz_tcp.cuh (packed-operand GEMM of IQN's 2048-row layers).
// (Round 1's register-loader kernel that lived here was retired in round 2: slower than the fp32-FMA kernels it was meant
// to replace — profiles/r01_tc_vs_simt.md — and superseded by the TMA-fed family.)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dz {
namespace tc {

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

// hi = rn_tf32(x), lo = rn_tf32(x - hi): round-to-nearest on both keeps the split error zero-mean
// (a truncating split biases every product towards zero).
__device__ __forceinline__ float rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
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

}  // namespace tc
}  // namespace dz
