// Torso (conv1-3) and 3136 -> 512 layer(s) of the batch-32 learner step on the TMA-fed tcgen05 kernels.
//   networks.py:181-204 dqn_torso, :82-103 conv, :207-221 dqn_value_head, :137-178 noisy_linear, :224-261 rainbow
// Data flow (all activations are stored once as tf32 hi/lo pairs by the producing epilogue, plus an fp32 copy
// for the layers that are still on the FMA kernels):
//
//   rows (uint8, replay store) --bulk copy--> smem --convert--> conv1 MMA --> act1 {hi,lo,f32} [P*B][h1][w1][32]
//   act1 --TMA im2col boxes--> conv2 MMA --> act2 [P*B][h2][w2][64] --TMA--> conv3 MMA --> act3 = features [P*B][feat]
//   W (fp32, [feat][512]) --TMA--> smem --in-place hi/lo split--> MN-major MMA x act3 --> split-K partials --> finish --> h1
//   dh1 --> W (K-major) MMA --> partials --> finish (ReLU mask) --> dact3 --TMA (zero-filled halo)--> conv3 dgrad --> dact2
//        --> conv2 dgrad (4 stride-parity classes) --> dact1
#include <algorithm>

#include "dz_umma_net.cuh"

namespace dz {

struct UmNet {
  UmNetDesc d;
  int h1, w1, h2, w2, h3, w3, feat, PB;
  int njt_fc;
  UmPlan plan;
  // activations / gradients
  float *act_hi[3], *act_lo[3], *act_f32[3];     // layer 1..3, all passes stacked
  float *dact_hi[3], *dact_lo[3], *dact_f32[3];  // layer 1..3 (pass 0)
  float *h1_buf, *dh1_f32, *dh1_hi, *dh1_lo;
  float *fc_part, *fcd_part;
  int fc_splits, fcd_splits, fc_nprob, fcd_nsrc;
  // conv weight images: [blob][layer] K-major [N][K]; dgrad images (online)
  float *wf_hi[2][3], *wf_lo[2][3], *wd3_hi, *wd3_lo, *wd2_hi, *wd2_lo;
  int map_wf1[2][2];                              // [blob][hi/lo] for the conv1 kernel
  UmLaunch l_conv2, l_conv3, l_dconv3, l_dconv2, l_fc, l_fcd, l_wconv3, l_wconv2;
  float *wg3_part, *wg2_part, *wg1_part;   // conv3 / conv2 weight-gradient split partials [S][K][64]; conv1: one [256][32] per CTA
  int wg3_splits, wg2_splits, wg1_ctas;
  int map_g1[2];                           // dact1 hi / lo as a flat [B*h1*w1][32] tensor, 32-byte-atom swizzle
  float* wg_scratch;                       // bias-gradient chunk sums [3][kWgMaxChunks][64]
  unsigned int* wg_ticket;                 // [4]
  int conv1_stag_bytes, conv1_tiles_per_pass;
  // noise-dependent problem fields (patched when the caller's noise buffer moves)
  struct Patch { int prob; int field; int64_t off; };   // field 0: A.scale_r, 2: A.scale_i, 1: scale_i (epilogue)
  std::vector<Patch> patches;
  const float* noise_cached = nullptr;
  std::string trace_tag;                   // debug: the launch with this tag writes CTA 0's clock stamps to trace_ptr
  long long* trace_ptr = nullptr;
  long long* tr(const char* tag) const { return trace_ptr && trace_tag == tag ? trace_ptr : nullptr; }
};

namespace {

using namespace um;

#define DZ_TRY_RC(expr) do { int _s = (expr); if (_s != DZ_OK) return _s; } while (0)

inline int conv_out_dim(int n, int k, int s) { return (n - k) / s + 1; }

// ------------------------------------------------------------------------------------------------
// Conv weight images: K-major [N][K] tf32 hi/lo for the forward GEMMs (conv1 carries the 1/255 of networks.py:193),
// and the input-gradient arrangements  Wd3[c][(kh,kw,n)] = W3[kh,kw,c,n],  Wd2[py,px][c][(ay,ax,n)] = W2[py+2ay,px+2ax,c,n].
// ------------------------------------------------------------------------------------------------
struct PackArgs {
  const float* blob[2];
  int64_t off_w[3];
  float* wf_hi[2][3]; float* wf_lo[2][3];
  float *wd3_hi, *wd3_lo, *wd2_hi, *wd2_lo;
};

__global__ void __launch_bounds__(256) um_pack_conv_kernel(const __grid_constant__ PackArgs a) {
  dz::pdl_enter();
  constexpr int kN[3] = {32, 64, 64}, kK[3] = {256, 512, 576};
  constexpr int kFwd = 32 * 256 + 64 * 512 + 64 * 576;   // 77824 per blob
  int e = blockIdx.x * 256 + threadIdx.x;
  float v;
  float *hi, *lo;
  if (e < 2 * kFwd) {
    const int b = e / kFwd;
    int r = e - b * kFwd;
    int L = 0;
    if (r >= kN[0] * kK[0]) { r -= kN[0] * kK[0]; L = 1; if (r >= kN[1] * kK[1]) { r -= kN[1] * kK[1]; L = 2; } }
    const int n = r / kK[L], k = r - n * kK[L];
    v = a.blob[b][a.off_w[L] + (int64_t)k * kN[L] + n];
    if (L == 0) v *= 0.0039215688593685627f;   // fl32(1/255): raw bytes are the (exact) MMA operand
    hi = a.wf_hi[b][L] + r; lo = a.wf_lo[b][L] + r;
  } else {
    e -= 2 * kFwd;
    if (e < 64 * 576) {                         // Wd3[c][(kh*3+kw)*64 + n]
      const int c = e / 576, r = e - c * 576, t = r >> 6, n = r & 63;
      v = a.blob[0][a.off_w[2] + ((int64_t)t * 64 + c) * 64 + n];
      hi = a.wd3_hi + e; lo = a.wd3_lo + e;
    } else {
      e -= 64 * 576;
      if (e >= 4 * 32 * 256) return;            // Wd2[(py*2+px)*32 + c][(ay*2+ax)*64 + n]
      const int row = e >> 8, r = e & 255, cls = row >> 5, c = row & 31, py = cls >> 1, px = cls & 1;
      const int t = r >> 6, n = r & 63, ay = t >> 1, ax = t & 1;
      const int kh = py + 2 * ay, kw = px + 2 * ax;
      v = a.blob[0][a.off_w[1] + ((int64_t)(kh * 4 + kw) * 32 + c) * 64 + n];
      hi = a.wd2_hi + e; lo = a.wd2_lo + e;
    }
  }
  const float h = rn_tf32(v);
  *hi = h;
  *lo = rn_tf32(v - h);
}

// ------------------------------------------------------------------------------------------------
// conv1: 8x8 stride 4 over the sampled uint8 observations, read IN PLACE from the replay store (the gather of
// replay.py:718-722 is this kernel's operand load).  Per 128-pixel output tile: one thread bulk-copies the
// contiguous input rows the tile needs (cp.async.bulk, <= 4 segments) into a double-buffered staging area;
// eight converter warps expand the bytes to exact tf32 values in the swizzled K-major A tile (K = 256 = 8 kernel
// rows x 32); one thread issues the MMAs against the resident weight image (hi/lo); four warps drain TMEM,
// add the bias, ReLU and write act1 as tf32 hi/lo (+ fp32).
// ------------------------------------------------------------------------------------------------
struct Conv1Args {
  CUtensorMap wmap[2][2];          // weight image [blob][hi / lo] (grid-constant: descriptor fetch from the constant bank)
  const uint8_t* const* rows[3];
  const float* bias[3];
  int blob[3];
  float *out_hi, *out_lo, *out_f32;
  int npass, B, W, oh, ow, m_pass, tiles_per_pass, ntiles, stag_bytes;
  long long* trace;                // debug: clock stamps of CTA 0
};

constexpr int kC1W = 65536, kC1A = 131072, kC1Epi = 4096;

__global__ void __launch_bounds__(kThreadsU, 1) conv1_umma_kernel(const __grid_constant__ Conv1Args a) {
  if (threadIdx.x < 4) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.wmap[threadIdx.x >> 1][threadIdx.x & 1])) : "memory");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* raw_full = reinterpret_cast<uint64_t*>(smem);   // [2] staged input rows landed
  uint64_t* raw_empty = raw_full + 2;                        // [2] converters are done reading them
  uint64_t* w_full = raw_empty + 2;                          // [1] weight image landed (only reloaded when the pass changes)
  uint64_t* a_ready = w_full + 1;                            // [4] kernel-row pair (2j, 2j+1) of the A tile converted
  uint64_t* a_empty = a_ready + 4;                           // [4] ... consumed by the MMAs
  uint64_t* acc_full = a_empty + 4;                          // [2]
  uint64_t* acc_empty = acc_full + 2;                        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint8_t* w_smem = smem + 1024;
  uint8_t* a_smem = w_smem + kC1W;
  uint8_t* stag = a_smem + kC1A;
  uint8_t* epi_smem = stag + 2 * (size_t)a.stag_bytes;       // 4 x 1 KB transposition patches of the epilogue warps

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) {
    if (lane == 0) {
      for (int b = 0; b < 2; ++b) { mbar_init(&raw_full[b], 1); mbar_init(&raw_empty[b], kConvWarps); mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
      for (int j = 0; j < 4; ++j) { mbar_init(&a_ready[j], kConvWarps); mbar_init(&a_empty[j], 1); }
      mbar_init(w_full, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  dz::pdl_enter();                        // set-up above overlaps the previous kernel's tail; data accesses start here
  const bool tr = a.trace != nullptr && blockIdx.x == 0;
  if (tr && threadIdx.x == 0) { a.trace[323] = clock64(); a.trace[324] = clock64(); }

  const int px = a.oh * a.ow;             // output pixels per image
  const int row_bytes = a.W * 4;          // one input row of 4-channel pixels
  // consecutive tiles per CTA: the weight image is reloaded only when the pass (online / target parameters) changes
  const int t_begin = (int)(((long long)blockIdx.x * a.ntiles) / gridDim.x);
  const int t_end = (int)(((long long)(blockIdx.x + 1) * a.ntiles) / gridDim.x);

  if (warp == 0) {
    // ---------------------------------------------------------------- producer
    int cur_pass = -1;
    for (int tile = t_begin, n = 0; tile < t_end; ++tile, ++n) {
      const int buf = n & 1;
      const int pass = tile / a.tiles_per_pass;
      const int m0 = (tile - pass * a.tiles_per_pass) * 128, m1 = min(m0 + 128, a.m_pass);
      mbar_wait(&raw_empty[buf], (((uint32_t)n >> 1) & 1u) ^ 1u);
      if (elect_one()) {
        const int b0 = m0 / px, b1 = (m1 - 1) / px;
        uint32_t total = 0;
        for (int b = b0; b <= b1; ++b) {
          const int plo = max(m0, b * px) - b * px, phi = min(m1, (b + 1) * px) - b * px;
          total += (uint32_t)((4 * ((phi - 1) / a.ow - plo / a.ow) + 8) * row_bytes);
        }
        mbar_expect_tx(&raw_full[buf], total);
        uint32_t off = 0;
        const uint32_t dst0 = smem_u32(stag + (size_t)buf * a.stag_bytes);
        for (int b = b0; b <= b1; ++b) {
          const int plo = max(m0, b * px) - b * px, phi = min(m1, (b + 1) * px) - b * px;
          const int oy0 = plo / a.ow;
          const uint32_t bytes = (uint32_t)((4 * ((phi - 1) / a.ow - oy0) + 8) * row_bytes);
          bulk_g2s(dst0 + off, a.rows[pass][b] + (size_t)(4 * oy0) * row_bytes, bytes, &raw_full[buf]);
          off += bytes;
        }
        if (tr && n < 64) a.trace[n] = clock64();                                                   // [0,64): row copies issued
      }
      __syncwarp();
      if (pass != cur_pass) {
        if (n > 0) mbar_wait(&a_empty[3], ((uint32_t)(n - 1)) & 1u);   // previous tile's MMAs are done with the old image
        if (elect_one()) {
          mbar_expect_tx(w_full, (uint32_t)kC1W);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int part = q >> 3, s = q & 7;
            tma_load_5d(smem_u32(w_smem) + part * 32768 + s * 4096, &a.wmap[a.blob[pass]][part], w_full, 32 * s, 0, 0, 0, 0);
          }
        }
        cur_pass = pass;
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc = make_idesc(128, 32, 0, 0);
    int cur_pass = -1, wn = 0;
    for (int tile = t_begin, n = 0; tile < t_end; ++tile, ++n) {
      const int pass = tile / a.tiles_per_pass;
      if (pass != cur_pass) { mbar_wait(w_full, (uint32_t)wn & 1u); ++wn; cur_pass = pass; }
      for (int slab = 0; slab < 8; ++slab) {
        if ((slab & 1) == 0) mbar_wait(&a_ready[slab >> 1], (uint32_t)n & 1u);
        const int g = n * 8 + slab, buf = g & 1;
        mbar_wait(&acc_empty[buf], (((uint32_t)g >> 1) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          // descriptor words: upper = SBO 1024 | version | SWIZZLE_128B, lower = LBO 16 | address >> 4 (advanced with adds)
          constexpr uint32_t up = (1024u >> 4) | (1u << 14) | (2u << 29);
          uint32_t al = (1u << 16) + ((smem_u32(a_smem) + slab * 16384) >> 4);
          uint32_t wh = (1u << 16) + ((smem_u32(w_smem) + slab * 4096) >> 4), wl = wh + (32768u >> 4);
          const uint32_t d = tmem_base + (uint32_t)(buf * 32);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = ((uint64_t)up << 32) | al;
            mma_tf32(d, da, ((uint64_t)up << 32) | wl, idesc, k > 0 ? 1u : 0u);
            mma_tf32(d, da, ((uint64_t)up << 32) | wh, idesc, 1u);
            al += 2; wh += 2; wl += 2;
          }
          mma_commit(&acc_full[buf]);
          if (slab & 1) mma_commit(&a_empty[slab >> 1]);
          if (slab == 7 && tile == t_end - 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        }
        __syncwarp();
      }
      if (tr && lane == 0 && n < 64) a.trace[128 + n] = clock64();                                  // [128,192): tile's MMAs issued
    }
    if (t_begin >= t_end && elect_one()) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  } else if (warp < 6) {
    // ---------------------------------------------------------------- epilogue
    const int quarter = warp & 3;
    for (int tile = t_begin, n = 0; tile < t_end; ++tile, ++n) {
      const int pass = tile / a.tiles_per_pass;
      const int m0 = (tile - pass * a.tiles_per_pass) * 128, m1 = min(m0 + 128, a.m_pass);
      float sum[32];
      {   // the bias is the initial value of the row sums (loaded while the tile's first MMAs are in flight)
        const float* __restrict__ bias = a.bias[pass];
#pragma unroll
        for (int t = 0; t < 32; t += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias + t);
          sum[t] = b4.x; sum[t + 1] = b4.y; sum[t + 2] = b4.z; sum[t + 3] = b4.w;
        }
      }
      for (int slab = 0; slab < 8; ++slab) {
        const int g = n * 8 + slab, buf = g & 1;
        mbar_wait(&acc_full[buf], ((uint32_t)g >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 32), r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int t = 0; t < 32; ++t) sum[t] += __uint_as_float(r[t]);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
      if (tr && warp == 2 && lane == 0 && n < 64) a.trace[192 + n] = clock64();                     // [192,256): tile drained
      // The warp's 32 rows x 32 channels are ONE contiguous 4 KB block of act1: transpose it, 8 channels at a time, through
      // a private XOR-swizzled 1 KB patch of shared memory (all that is left next to the 128 KB A tile), so that a store
      // instruction writes 16 rows x 32 contiguous bytes (full sectors) instead of 32 rows x 16 bytes.
      float4* patch = reinterpret_cast<float4*>(epi_smem + (warp - 2) * 1024);
      const int mw0 = m0 + quarter * 32;                       // first row of this warp's block
      const long long dst0 = ((long long)pass * a.m_pass + mw0) * 32;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int t = q * 8 + cc * 4;
          patch[lane * 2 + (cc ^ ((lane >> 2) & 1))] = make_float4(fmaxf(sum[t], 0.f), fmaxf(sum[t + 1], 0.f), fmaxf(sum[t + 2], 0.f), fmaxf(sum[t + 3], 0.f));
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int row = j * 16 + (lane >> 1), cc = lane & 1;
          if (mw0 + row < m1) {
            const float4 v = patch[row * 2 + (cc ^ ((row >> 2) & 1))];
            float4 h, l;
            split4(v, h, l);
            *reinterpret_cast<float4*>(a.out_hi + dst0 + row * 32 + q * 8 + 4 * cc) = h;
            *reinterpret_cast<float4*>(a.out_lo + dst0 + row * 32 + q * 8 + 4 * cc) = l;
          }
        }
        __syncwarp();
      }
      if (tr && warp == 2 && lane == 0 && n < 64) a.trace[256 + n] = clock64();                     // [256,320): tile stored
    }
  } else {
    // ---------------------------------------------------------------- converters: uint8 rows -> exact tf32 A tile
    const int ct = threadIdx.x - 6 * 32;
    const int r = ct & 127, khp = ct >> 7;
    for (int tile = t_begin, n = 0; tile < t_end; ++tile, ++n) {
      const int buf = n & 1;
      const int pass = tile / a.tiles_per_pass;
      const int m0 = (tile - pass * a.tiles_per_pass) * 128, m1 = min(m0 + 128, a.m_pass);
      const int m = m0 + r;
      const bool valid = m < m1;
      // staging offset of this row's patch origin
      int src_off = 0;
      if (valid) {
        const int b0 = m0 / px, b = m / px, p = m - b * px, oy = p / a.ow, ox = p - oy * a.ow;
        int off = 0;
        for (int bb = b0; bb < b; ++bb) {
          const int plo = max(m0, bb * px) - bb * px, phi = min(m1, (bb + 1) * px) - bb * px;
          off += (4 * ((phi - 1) / a.ow - plo / a.ow) + 8) * row_bytes;
        }
        const int plo = max(m0, b * px) - b * px;
        src_off = off + (4 * (oy - plo / a.ow)) * row_bytes + 16 * ox;
      }
      mbar_wait(&raw_full[buf], ((uint32_t)n >> 1) & 1u);
      if (tr && ct == 0 && n < 64) a.trace[64 + n] = clock64();                                     // [64,128): input rows landed
      const uint8_t* src = stag + (size_t)buf * a.stag_bytes + src_off;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int kh = 2 * it + khp;
        uint4 u[2];
        if (valid) {
          u[0] = *reinterpret_cast<const uint4*>(src + kh * row_bytes);
          u[1] = *reinterpret_cast<const uint4*>(src + kh * row_bytes + 16);
        } else {
          u[0] = make_uint4(0, 0, 0, 0); u[1] = u[0];
        }
        mbar_wait(&a_empty[it], ((uint32_t)n & 1u) ^ 1u);     // the previous tile's MMAs have consumed this kernel-row pair
        uint8_t* dstrow = a_smem + kh * 16384 + r * 128;
        const uint32_t w[8] = {u[0].x, u[0].y, u[0].z, u[0].w, u[1].x, u[1].y, u[1].z, u[1].w};
#pragma unroll
        for (int kw = 0; kw < 8; ++kw) {
          // 0x4B000000 | byte = 8388608 + byte exactly; subtracting 2^23 leaves the byte as a float (an exact tf32 number)
          float4 f;
          f.x = __uint_as_float(__byte_perm(w[kw], 0x4B000000u, 0x7440)) - 8388608.0f;
          f.y = __uint_as_float(__byte_perm(w[kw], 0x4B000000u, 0x7441)) - 8388608.0f;
          f.z = __uint_as_float(__byte_perm(w[kw], 0x4B000000u, 0x7442)) - 8388608.0f;
          f.w = __uint_as_float(__byte_perm(w[kw], 0x4B000000u, 0x7443)) - 8388608.0f;
          *reinterpret_cast<float4*>(dstrow + ((kw ^ (r & 7)) << 4)) = f;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[it]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&raw_empty[buf]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64) : "memory");
  }
  if (tr && threadIdx.x == 0) a.trace[322] = clock64();
}

// ------------------------------------------------------------------------------------------------
// conv1 weight gradient: dW1[k][n] = (1/255) * sum_m X[m][k] * dact1[m][n], reduction over the B * h1 * w1 output pixels of
// pass 0.  Same staging as the forward kernel (bulk-copied uint8 rows -> exact tf32 A tile), but the tile is written in
// the 32-byte-atom swizzle and consumed through MN-major descriptors (rows = reduction index): A^T is never formed.
// G = dact1 hi/lo arrives by TMA.  Two accumulators (k 0..127 | 128..255) x two TMEM buffers; one partial per CTA.
// ------------------------------------------------------------------------------------------------
struct Conv1WgArgs {
  CUtensorMap gmap[2];             // dact1 hi / lo
  const uint8_t* const* rows;
  float* partial;                // [gridDim.x][256][32]
  int B, W, oh, ow, m_pass, ntiles, stag_bytes;
};

constexpr int kC1G = 32768;      // dact1 tile: hi | lo, [128 rows][32 n] each

__global__ void __launch_bounds__(kThreadsU, 1) conv1_wgrad_umma_kernel(const __grid_constant__ Conv1WgArgs a) {
  if (threadIdx.x < 2) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.gmap[threadIdx.x])) : "memory");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* raw_full = reinterpret_cast<uint64_t*>(smem);   // [2]
  uint64_t* raw_empty = raw_full + 2;                        // [2]
  uint64_t* g_full = raw_empty + 2;                          // [1] dact1 tile landed
  uint64_t* a_ready = g_full + 1;                            // [4] kernel-row pairs converted
  uint64_t* t_done = a_ready + 4;                            // [1] all MMAs of the tile done (A and G tiles free)
  uint64_t* acc_full = t_done + 1;                           // [2]
  uint64_t* acc_empty = acc_full + 2;                        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  uint8_t* g_smem = smem + 1024;
  uint8_t* a_smem = g_smem + kC1G;
  uint8_t* stag = a_smem + kC1A;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) {
    if (lane == 0) {
      for (int b = 0; b < 2; ++b) { mbar_init(&raw_full[b], 1); mbar_init(&raw_empty[b], kConvWarps); mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
      for (int j = 0; j < 4; ++j) mbar_init(&a_ready[j], kConvWarps);
      mbar_init(g_full, 1); mbar_init(t_done, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  dz::pdl_enter();                        // set-up above overlaps the previous kernel's tail; data accesses start here

  const int px = a.oh * a.ow;
  const int row_bytes = a.W * 4;
  const int t_begin = (int)(((long long)blockIdx.x * a.ntiles) / gridDim.x);
  const int t_end = (int)(((long long)(blockIdx.x + 1) * a.ntiles) / gridDim.x);

  if (warp == 0) {
    // ---------------------------------------------------------------- producer
    for (int tile = t_begin, n = 0; tile < t_end; ++tile, ++n) {
      const int buf = n & 1;
      const int m0 = tile * 128, m1 = min(m0 + 128, a.m_pass);
      mbar_wait(&raw_empty[buf], (((uint32_t)n >> 1) & 1u) ^ 1u);
      if (elect_one()) {
        const int b0 = m0 / px, b1 = (m1 - 1) / px;
        uint32_t total = 0;
        for (int b = b0; b <= b1; ++b) {
          const int plo = max(m0, b * px) - b * px, phi = min(m1, (b + 1) * px) - b * px;
          total += (uint32_t)((4 * ((phi - 1) / a.ow - plo / a.ow) + 8) * row_bytes);
        }
        mbar_expect_tx(&raw_full[buf], total);
        uint32_t off = 0;
        const uint32_t dst0 = smem_u32(stag + (size_t)buf * a.stag_bytes);
        for (int b = b0; b <= b1; ++b) {
          const int plo = max(m0, b * px) - b * px, phi = min(m1, (b + 1) * px) - b * px;
          const int oy0 = plo / a.ow;
          const uint32_t bytes = (uint32_t)((4 * ((phi - 1) / a.ow - oy0) + 8) * row_bytes);
          bulk_g2s(dst0 + off, a.rows[b] + (size_t)(4 * oy0) * row_bytes, bytes, &raw_full[buf]);
          off += bytes;
        }
      }
      __syncwarp();
      if (n > 0) mbar_wait(t_done, ((uint32_t)(n - 1)) & 1u);
      if (elect_one()) {
        mbar_expect_tx(g_full, (uint32_t)kC1G);
        tma_load_5d(smem_u32(g_smem), &a.gmap[0], g_full, 0, m0, 0, 0, 0);
        tma_load_5d(smem_u32(g_smem) + 16384, &a.gmap[1], g_full, 0, m0, 0, 0, 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    const uint32_t idesc = make_idesc(128, 32, 1, 1);     // both operands MN-major (rows of the tiles are the reduction)
    for (int tile = t_begin, n = 0; tile < t_end; ++tile, ++n) {
      mbar_wait(g_full, (uint32_t)n & 1u);
      for (int mt = 0; mt < 2; ++mt) {                      // k 0..127 (kernel rows 0-3) | k 128..255 (kernel rows 4-7)
        mbar_wait(&a_ready[2 * mt], (uint32_t)n & 1u);
        mbar_wait(&a_ready[2 * mt + 1], (uint32_t)n & 1u);
        const int g = n * 2 + mt, buf = g & 1;
        mbar_wait(&acc_empty[buf], (((uint32_t)g >> 1) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          // descriptor words: upper = SBO 512 | version | 32-byte-atom swizzle, lower = LBO 16384 | address >> 4
          constexpr uint32_t up = (512u >> 4) | (1u << 14) | (1u << 29);
          constexpr uint32_t lbo = (16384u >> 4) << 16;
          uint32_t al = lbo + ((smem_u32(a_smem) + mt * 4 * 16384) >> 4);
          uint32_t gh = lbo + (smem_u32(g_smem) >> 4), gl = gh + (16384u >> 4);
          const uint32_t d = tmem_base + (uint32_t)(buf * 64 + mt * 32);
#pragma unroll 4
          for (int k = 0; k < 16; ++k) {
            const uint64_t da = ((uint64_t)up << 32) | al;
            mma_tf32(d, da, ((uint64_t)up << 32) | gl, idesc, k > 0 ? 1u : 0u);
            mma_tf32(d, da, ((uint64_t)up << 32) | gh, idesc, 1u);
            al += 64; gh += 64; gl += 64;
          }
          mma_commit(&acc_full[buf]);
          if (mt == 1) mma_commit(t_done);
          if (mt == 1 && tile == t_end - 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ---------------------------------------------------------------- epilogue: rows k of both accumulators
    const int quarter = warp & 3;
    float sum[2][32];
#pragma unroll
    for (int t = 0; t < 32; ++t) { sum[0][t] = 0.f; sum[1][t] = 0.f; }
    for (int tile = t_begin, n = 0; tile < t_end; ++tile, ++n) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int g = n * 2 + mt, buf = g & 1;
        mbar_wait(&acc_full[buf], ((uint32_t)g >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * 64 + mt * 32), r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int t = 0; t < 32; ++t) sum[mt][t] += __uint_as_float(r[t]);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[buf]);
      }
    }
    const float inv255 = 0.0039215688593685627f;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float* dst = a.partial + ((long long)blockIdx.x * 256 + mt * 128 + quarter * 32 + lane) * 32;
#pragma unroll
      for (int t = 0; t < 32; t += 4)
        *reinterpret_cast<float4*>(dst + t) = make_float4(sum[mt][t] * inv255, sum[mt][t + 1] * inv255, sum[mt][t + 2] * inv255, sum[mt][t + 3] * inv255);
    }
  } else {
    // ---------------------------------------------------------------- converters
    const int ct = threadIdx.x - 6 * 32;
    const int r = ct & 127, khp = ct >> 7;
    for (int tile = t_begin, n = 0; tile < t_end; ++tile, ++n) {
      const int buf = n & 1;
      const int m0 = tile * 128, m1 = min(m0 + 128, a.m_pass);
      const int m = m0 + r;
      const bool valid = m < m1;
      int src_off = 0;
      if (valid) {
        const int b0 = m0 / px, b = m / px, p = m - b * px, oy = p / a.ow, ox = p - oy * a.ow;
        int off = 0;
        for (int bb = b0; bb < b; ++bb) {
          const int plo = max(m0, bb * px) - bb * px, phi = min(m1, (bb + 1) * px) - bb * px;
          off += (4 * ((phi - 1) / a.ow - plo / a.ow) + 8) * row_bytes;
        }
        const int plo = max(m0, b * px) - b * px;
        src_off = off + (4 * (oy - plo / a.ow)) * row_bytes + 16 * ox;
      }
      mbar_wait(&raw_full[buf], ((uint32_t)n >> 1) & 1u);
      if (n > 0) mbar_wait(t_done, ((uint32_t)(n - 1)) & 1u);     // previous tile's MMAs have consumed the A tile
      const uint8_t* src = stag + (size_t)buf * a.stag_bytes + src_off;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int kh = 2 * it + khp;
        uint4 u[2];
        if (valid) {
          u[0] = *reinterpret_cast<const uint4*>(src + kh * row_bytes);
          u[1] = *reinterpret_cast<const uint4*>(src + kh * row_bytes + 16);
        } else {
          u[0] = make_uint4(0, 0, 0, 0); u[1] = u[0];     // rows beyond the batch contribute zeros to the reduction
        }
        uint8_t* dstrow = a_smem + kh * 16384 + r * 128;
        const uint32_t w[8] = {u[0].x, u[0].y, u[0].z, u[0].w, u[1].x, u[1].y, u[1].z, u[1].w};
#pragma unroll
        for (int kw = 0; kw < 8; ++kw) {
          float4 f;
          f.x = __uint_as_float(__byte_perm(w[kw], 0x4B000000u, 0x7440)) - 8388608.0f;
          f.y = __uint_as_float(__byte_perm(w[kw], 0x4B000000u, 0x7441)) - 8388608.0f;
          f.z = __uint_as_float(__byte_perm(w[kw], 0x4B000000u, 0x7442)) - 8388608.0f;
          f.w = __uint_as_float(__byte_perm(w[kw], 0x4B000000u, 0x7443)) - 8388608.0f;
          // 128B swizzle with 32-byte atoms: the 32-byte chunk index is XOR-ed with (row & 3)
          *reinterpret_cast<float4*>(dstrow + ((((kw >> 1) ^ (r & 3)) << 5) | ((kw & 1) << 4))) = f;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_ready[it]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&raw_empty[buf]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Finish kernels of the split FC GEMMs
// ------------------------------------------------------------------------------------------------

// h1[pass][stream][m][n] = relu( sum_s P + b_mu[n] + eps_out[n] * b_sigma[n] )   (networks.py:160-178; P already holds x(Wmu + Wsigma*noise))
struct FcFinishArgs {
  const float* part; int S, B, nstream, noisy, npass;
  const float* bmu[3][2]; const float* bsig[3][2]; const float* eps_out[3][2];
  float* h1;
};

__global__ void __launch_bounds__(256) um_fc_finish_kernel(const __grid_constant__ FcFinishArgs a) {
  dz::pdl_enter();
  const int ps = blockIdx.y, pass = ps / a.nstream, st = ps - pass * a.nstream;
  const int total4 = a.B * 128;
  const long long pstride = (long long)a.B * 512;
  const float* pm = a.part + (long long)ps * a.S * pstride;
  for (int i4 = blockIdx.x * 256 + threadIdx.x; i4 < total4; i4 += gridDim.x * 256) {
    const int i = i4 << 2, n = i & 511;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < a.S; ++s) {
      const float4 x = *reinterpret_cast<const float4*>(pm + s * pstride + i);
      v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
    }
    const float4 b = *reinterpret_cast<const float4*>(a.bmu[pass][st] + n);
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    if (a.noisy) {     // the weights' noise went into the GEMM operand; the bias noise is b_sigma * eps_out (networks.py:172-177)
      const float4 bs = *reinterpret_cast<const float4*>(a.bsig[pass][st] + n);
      const float4 eo = *reinterpret_cast<const float4*>(a.eps_out[pass][st] + n);
      v.x = fmaf(bs.x, eo.x, v.x); v.y = fmaf(bs.y, eo.y, v.y); v.z = fmaf(bs.z, eo.z, v.z); v.w = fmaf(bs.w, eo.w, v.w);
    }
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    *reinterpret_cast<float4*>(a.h1 + (long long)ps * pstride + i) = v;
  }
}

// dact3[m][k] = [act3 > 0] * sum over sources and splits of the partials; fp32 + tf32 hi/lo.
__global__ void __launch_bounds__(256) um_fcd_finish_kernel(const float* __restrict__ part, int nparts, long long stride,
                                                            const float* __restrict__ act_hi, float* __restrict__ out,
                                                            float* __restrict__ out_hi, float* __restrict__ out_lo, long long total4) {
  dz::pdl_enter();
  for (long long i4 = blockIdx.x * 256LL + threadIdx.x; i4 < total4; i4 += (long long)gridDim.x * 256) {
    const long long i = i4 << 2;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < nparts; ++s) {
      const float4 x = *reinterpret_cast<const float4*>(part + s * stride + i);
      v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
    }
    const float4 m = *reinterpret_cast<const float4*>(act_hi + i);
    v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
    float4 h, l;
    split4(v, h, l);
    *reinterpret_cast<float4*>(out + i) = v;
    *reinterpret_cast<float4*>(out_hi + i) = h;
    *reinterpret_cast<float4*>(out_lo + i) = l;
  }
}


// Weight-gradient finish: dW[k][n] = sum_s partial[s][k][n] (fixed order), and the bias gradient db[n] = sum_m G[m][n]
// as a deterministic two-level column sum of the (fp32) output gradient: 128-row chunks in parallel, the last block to
// finish adds the chunk sums in chunk order.  One launch covers several layers: blockIdx.y = layer; blocks
// [0, kWgSumBlocks) add the partials, blocks beyond do the bias chunks.
// Partial sums: a block owns 32 consecutive float4 columns; its 8 thread groups stride over the S partials (so the
// up-to-148 loads of one column are 8 independent chains instead of one dependent chain), then group sums are added in
// group order through shared memory — a fixed association, hence bit-deterministic.
constexpr int kWgChunkRows = 128, kWgMaxChunks = 256, kWgGroups = 8;
// norm_part: this layer's slots of the split global gradient norm: [0, sum_blocks) = sum of squares of the dW values each
// sum block wrote, [sum_blocks] = sum of squares of db (dz_learner.cu: split_norm).
struct WgFinish { const float* partial; int S; long long stride; int KN; float* dW; const float* G; int M, N; float* db; float* scratch;
                  unsigned int* ticket; int sum_blocks; float* norm_part; };

__global__ void __launch_bounds__(256) um_wgrad_finish_kernel(const __grid_constant__ WgFinish f) {
  dz::pdl_enter();
  __shared__ float4 red4[256];
  __shared__ bool last;
  if ((int)blockIdx.x >= f.sum_blocks) {
    float* red = reinterpret_cast<float*>(red4);
    const int chunk = blockIdx.x - f.sum_blocks, nchunks = (f.M + kWgChunkRows - 1) / kWgChunkRows;
    if (!f.db || chunk >= nchunks) return;
    const int n = threadIdx.x % f.N, g = threadIdx.x / f.N, G = 256 / f.N;
    const int m1 = min(f.M, (chunk + 1) * kWgChunkRows);
    float acc = 0.f;
    for (int m = chunk * kWgChunkRows + g; m < m1; m += G) acc += f.G[(long long)m * f.N + n];
    red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x < f.N) {
      float t = 0.f;
      for (int q = 0; q < G; ++q) t += red[q * f.N + threadIdx.x];
      f.scratch[chunk * 64 + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(f.ticket, 1u) == (unsigned)(nchunks - 1);
    __syncthreads();
    if (last) {
      __threadfence();
      // chunk sums: G thread groups stride over the chunks (independent L2 loads), group sums added in group order
      float t = 0.f;
#pragma unroll 4
      for (int c = g; c < nchunks; c += G) t += __ldcg(f.scratch + c * 64 + n);
      red[threadIdx.x] = t;
      __syncthreads();
      float sq = 0.f;
      if (threadIdx.x < f.N) {
        float tt = 0.f;
        for (int q = 0; q < G; ++q) tt += red[q * f.N + threadIdx.x];
        f.db[threadIdx.x] = tt;
        sq = tt * tt;
      }
      __syncthreads();
      red[threadIdx.x] = sq;
      __syncthreads();
      if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int q = 0; q < f.N; ++q) tot += red[q];
        if (f.norm_part) f.norm_part[f.sum_blocks] = tot;
        *f.ticket = 0;
      }
    }
    return;
  }
  const int total4 = f.KN >> 2;
  const int col = blockIdx.x * 32 + (threadIdx.x & 31), g = threadIdx.x >> 5;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < total4) {
    const float* src = f.partial + ((long long)col << 2);
#pragma unroll 4
    for (int s = g; s < f.S; s += kWgGroups) {
      const float4 x = *reinterpret_cast<const float4*>(src + s * f.stride);
      v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
    }
  }
  red4[threadIdx.x] = v;
  __syncthreads();
  if (g == 0) {
    float sq = 0.f;
    if (col < total4) {
#pragma unroll
      for (int q = 1; q < kWgGroups; ++q) {
        const float4 x = red4[q * 32 + threadIdx.x];
        v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
      }
      *reinterpret_cast<float4*>(f.dW + ((long long)col << 2)) = v;
      sq = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
    }
    sq = warp_sum(sq);
    if (threadIdx.x == 0 && f.norm_part) f.norm_part[blockIdx.x] = sq;
  }
}

// L2 prefetch of the 3136 -> 512 weight matrices (both parameter sets): issued on the side stream at the start of the
// step, while the sampler and the conv stack keep HBM idle, so that the weight-streaming GEMM later reads L2, not DRAM.
struct PrefetchArgs { const float* ptr[8]; long long lines[8]; int n; };
__global__ void __launch_bounds__(256) um_prefetch_kernel(const __grid_constant__ PrefetchArgs a) {
  dz::pdl_enter();
  for (int t = 0; t < a.n; ++t) {
    const char* base = reinterpret_cast<const char*>(a.ptr[t]);
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < a.lines[t]; i += (long long)gridDim.x * 256)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + i * 128));
  }
}

}  // namespace

int um_split(const float* x, float* hi, float* lo, long long n, void* stream);   // dz_umma.cu

// ------------------------------------------------------------------------------------------------
// Plan
// ------------------------------------------------------------------------------------------------

bool um_net_supported(const UmNetDesc& d) {
  if (d.B < 1 || d.B > 64 || d.npass < 1 || d.npass > 3) return false;
  if (d.W % 4 || d.H < 36 || d.W < 36) return false;
  const int h1 = conv_out_dim(d.H, 8, 4), w1 = conv_out_dim(d.W, 8, 4);
  if ((h1 & 1) || (w1 & 1)) return false;              // stride-2 parity view of act1
  const int h2 = conv_out_dim(h1, 4, 2), w2 = conv_out_dim(w1, 4, 2);
  const int h3 = conv_out_dim(h2, 3, 1), w3 = conv_out_dim(w2, 3, 1);
  if (h3 < 1 || w3 < 1) return false;
  if (h2 * w2 > 128 || (h1 / 2) * (w1 / 2) > 128 || h3 * w3 > 128) return false;
  if (h1 * w1 < 64) return false;                      // a 128-pixel conv1 tile spans at most 3 images
  // conv1 staging: worst case bytes of one tile
  const int px = h1 * w1, m_pass = d.B * px;
  int worst = 0;
  for (int m0 = 0; m0 < m_pass; m0 += 128) {
    const int m1 = std::min(m0 + 128, m_pass);
    int tot = 0;
    for (int b = m0 / px; b <= (m1 - 1) / px; ++b) {
      const int plo = std::max(m0, b * px) - b * px, phi = std::min(m1, (b + 1) * px) - b * px;
      tot += (4 * ((phi - 1) / w1 - plo / w1) + 8) * d.W * 4;
    }
    worst = std::max(worst, tot);
  }
  if (2 * ((worst + 127) / 128 * 128) + 2048 + 65536 + 131072 + 4096 > 227 * 1024) return false;
  return true;
}

namespace {

struct Geo { int h1, w1, h2, w2, h3, w3, feat, PB; };
Geo geo_of(const UmNetDesc& d) {
  Geo g;
  g.h1 = conv_out_dim(d.H, 8, 4); g.w1 = conv_out_dim(d.W, 8, 4);
  g.h2 = conv_out_dim(g.h1, 4, 2); g.w2 = conv_out_dim(g.w1, 4, 2);
  g.h3 = conv_out_dim(g.h2, 3, 1); g.w3 = conv_out_dim(g.w2, 3, 1);
  g.feat = g.h3 * g.w3 * 64; g.PB = d.npass * d.B;
  return g;
}

struct Carver {
  char* base; int64_t used = 0;
  float* f(int64_t n) {
    int64_t bytes = (n * 4 + 255) / 256 * 256;
    float* p = base ? reinterpret_cast<float*>(base + used) : nullptr;
    used += bytes;
    return p;
  }
};

int fc_splits_for(int nprob, int nk) { return std::max(1, std::min(std::min(148 / (nprob * 4), 24), nk)); }

int64_t carve_net(UmNet* n, char* base) {
  const UmNetDesc& d = n->d;
  const Geo g = geo_of(d);
  n->h1 = g.h1; n->w1 = g.w1; n->h2 = g.h2; n->w2 = g.w2; n->h3 = g.h3; n->w3 = g.w3; n->feat = g.feat; n->PB = g.PB;
  Carver c{base};
  const int64_t a1 = (int64_t)g.PB * g.h1 * g.w1 * 32, a2 = (int64_t)g.PB * g.h2 * g.w2 * 64, a3 = (int64_t)g.PB * g.feat;
  const int64_t sz[3] = {a1, a2, a3};
  for (int L = 0; L < 3; ++L) { n->act_hi[L] = c.f(sz[L]); n->act_lo[L] = c.f(sz[L]); n->act_f32[L] = c.f(sz[L]); }
  const int64_t dz_[3] = {(int64_t)d.B * g.h1 * g.w1 * 32, (int64_t)d.B * g.h2 * g.w2 * 64, (int64_t)d.B * g.feat};
  for (int L = 0; L < 3; ++L) { n->dact_hi[L] = c.f(dz_[L]); n->dact_lo[L] = c.f(dz_[L]); n->dact_f32[L] = c.f(dz_[L]); }
  const int kN[3] = {32, 64, 64}, kK[3] = {256, 512, 576};
  for (int b = 0; b < 2; ++b)
    for (int L = 0; L < 3; ++L) { n->wf_hi[b][L] = c.f(kN[L] * kK[L]); n->wf_lo[b][L] = c.f(kN[L] * kK[L]); }
  n->wd3_hi = c.f(64 * 576); n->wd3_lo = c.f(64 * 576);
  n->wd2_hi = c.f(128 * 256); n->wd2_lo = c.f(128 * 256);
  {
    const int groups = (d.B + 7) / 8;
    n->wg3_splits = std::max(1, std::min(g.h3, 148 / 5));             // stages = h3 * groups, split over the output rows
    n->wg2_splits = std::max(1, std::min(g.h2, 148 / 8));
    (void)groups;
    n->wg3_part = c.f((int64_t)n->wg3_splits * 576 * 64);
    n->wg2_part = c.f((int64_t)n->wg2_splits * 512 * 64);
    n->wg1_ctas = std::min(148, (d.B * g.h1 * g.w1 + 127) / 128);
    n->wg1_part = c.f((int64_t)n->wg1_ctas * 256 * 32);
    n->wg_scratch = c.f(3 * 256 * 64);
    n->wg_ticket = reinterpret_cast<unsigned int*>(c.f(64));
  }
  n->h1_buf = nullptr; n->dh1_f32 = n->dh1_hi = n->dh1_lo = nullptr; n->fc_part = n->fcd_part = nullptr;
  n->fc_nprob = n->fcd_nsrc = 0; n->fc_splits = n->fcd_splits = 1;
  if (d.use_fc) {
    n->fc_nprob = d.npass * d.nstream;      // noisy layers: mu and sigma are combined in the converter (one GEMM per stream)
    n->fcd_nsrc = d.nstream;
    n->fc_splits = fc_splits_for(n->fc_nprob, g.feat / 32);
    const int ktiles = (g.feat + 127) / 128;
    n->fcd_splits = std::max(1, std::min(148 / (n->fcd_nsrc * ktiles), 8));
    n->h1_buf = c.f((int64_t)d.npass * d.nstream * d.B * 512);
    n->dh1_f32 = c.f((int64_t)d.nstream * d.B * 512);
    n->dh1_hi = c.f((int64_t)d.nstream * d.B * 512);
    n->dh1_lo = c.f((int64_t)d.nstream * d.B * 512);
    n->fc_part = c.f((int64_t)n->fc_nprob * n->fc_splits * d.B * 512);
    n->fcd_part = c.f((int64_t)n->fcd_nsrc * n->fcd_splits * d.B * g.feat);
  }
  return c.used;
}

void push_op(UmPlan& pl, int map, uint32_t off, int c0, int c1, int c2, int c3, int c4) {
  UmTmaOp o;
  memset(&o, 0, sizeof(o));
  o.map = (uint32_t)map; o.smem_off = off; o.c[0] = c0; o.c[1] = c1; o.c[2] = c2; o.c[3] = c3; o.c[4] = c4;
  pl.ops.push_back(o);
}

int build_plan(UmNet* n) {
  const UmNetDesc& d = n->d;
  UmPlan& pl = n->plan;
  const int B = d.B, PB = n->PB, h1 = n->h1, w1 = n->w1, h2 = n->h2, w2 = n->w2, h3 = n->h3, w3 = n->w3, feat = n->feat;
  const int kN[3] = {32, 64, 64}, kK[3] = {256, 512, 576};
#define MAP_OR_FAIL(var, ...) const int var = pl.add_map(__VA_ARGS__); if (var < 0) return DZ_EINVAL;

  // ---- weight image maps: [blob][layer][part]; box rows = N tile used by the layer
  int m_wf[2][3][2];
  const uint32_t wf_rows[3] = {32, 64, 32};   // conv1: N 32; conv2: full 64; conv3: halves of 32
  for (int b = 0; b < 2; ++b)
    for (int L = 0; L < 3; ++L)
      for (int part = 0; part < 2; ++part) {
        uint64_t dims[2] = {(uint64_t)kK[L], (uint64_t)kN[L]}, strides[1] = {(uint64_t)kK[L] * 4};
        uint32_t box[2] = {32, wf_rows[L]};
        m_wf[b][L][part] = pl.add_map(part ? n->wf_lo[b][L] : n->wf_hi[b][L], 2, dims, strides, box);
        if (m_wf[b][L][part] < 0) return DZ_EINVAL;
      }
  for (int b = 0; b < 2; ++b) { n->map_wf1[b][0] = m_wf[b][0][0]; n->map_wf1[b][1] = m_wf[b][0][1]; }

  // =========================================================================== conv2 forward
  {
    int m_a[2];
    for (int part = 0; part < 2; ++part) {
      uint64_t dims[5] = {64, (uint64_t)w1 / 2, 2, (uint64_t)h1 / 2, (uint64_t)PB};
      uint64_t strides[4] = {256, (uint64_t)w1 * 128, (uint64_t)2 * w1 * 128, (uint64_t)h1 * w1 * 128};
      uint32_t box[5] = {32, (uint32_t)w2, 1, (uint32_t)h2, 1};
      m_a[part] = pl.add_map(part ? n->act_lo[0] : n->act_hi[0], 5, dims, strides, box);
      if (m_a[part] < 0) return DZ_EINVAL;
    }
    const int rows = h2 * w2;
    UmOperand A = um_kmajor(rows, true, false), Bo = um_kmajor(64, true, false);
    const uint32_t a_bytes = A.part_bytes * 2;
    n->l_conv2.cta0 = (int)pl.ctas.size(); n->l_conv2.njt = 64; n->l_conv2.stage_bytes = a_bytes + Bo.part_bytes * 2;
    // the MMA reads 128 rows of every A part: keep the (garbage) tail rows inside the stage
    n->l_conv2.stage_bytes = std::max<uint32_t>(n->l_conv2.stage_bytes, A.part_bytes + 16384);
    n->l_conv2.stages = std::min<int>(kStagesMax, (int)((226 * 1024 - 1024 - kCtlBytes) / n->l_conv2.stage_bytes));
    for (int p = 0; p < d.npass; ++p) {
      const int blob = d.pass_target[p] ? 1 : 0;
      UmProblem pr;
      memset(&pr, 0, sizeof(pr));
      pr.A = A; pr.B = Bo; pr.ksteps = 4; pr.run_stages = 1; pr.red_per_stage = 32; pr.epi = UM_EPI_ROWS;
      pr.MI = rows; pr.NJ = 64;
      pr.out_hi = n->act_hi[1]; pr.out_lo = n->act_lo[1]; pr.out_f32 = nullptr; pr.out_ld = 64;
      pr.bias = (blob ? d.target : d.online) + d.off_conv_b[1]; pr.relu = 1;
      pr.pw = 1 << 20; pr.rs_outer = 0; pr.rs_inner = 1;
      const int prob = (int)pl.probs.size();
      pl.probs.push_back(pr);
      for (int b = 0; b < B; ++b) {
        const int img = p * B + b;
        UmCta c;
        memset(&c, 0, sizeof(c));
        c.prob = (uint32_t)prob; c.op0 = (uint32_t)pl.ops.size(); c.nstages = 16; c.ops_per_stage = 4;
        c.tx_bytes = (uint32_t)(2 * rows * 128 + 2 * 64 * 128);
        c.row_base = img * rows; c.ph_valid = 1; c.pw_valid = rows;
        for (int kh = 0; kh < 4; ++kh)
          for (int kw = 0; kw < 4; ++kw) {
            const int s = kh * 4 + kw;
            for (int part = 0; part < 2; ++part) push_op(pl, m_a[part], part * A.part_bytes, 32 * (kw & 1), kw >> 1, kh & 1, kh >> 1, img);
            for (int part = 0; part < 2; ++part) push_op(pl, m_wf[blob][1][part], a_bytes + part * Bo.part_bytes, 32 * s, 0, 0, 0, 0);
          }
        pl.ctas.push_back(c);
      }
    }
    n->l_conv2.nctas = (int)pl.ctas.size() - n->l_conv2.cta0;
  }

  // =========================================================================== conv3 forward (two 32-channel halves)
  {
    const int G = (B % 2 == 0 && 2 * h3 * w3 <= 128) ? 2 : 1;
    int m_a[2];
    for (int part = 0; part < 2; ++part) {
      uint64_t dims[4] = {64, (uint64_t)w2, (uint64_t)h2, (uint64_t)PB};
      uint64_t strides[3] = {256, (uint64_t)w2 * 256, (uint64_t)h2 * w2 * 256};
      uint32_t box[4] = {32, (uint32_t)w3, (uint32_t)h3, (uint32_t)G};
      m_a[part] = pl.add_map(part ? n->act_lo[1] : n->act_hi[1], 4, dims, strides, box);
      if (m_a[part] < 0) return DZ_EINVAL;
    }
    const int rows = G * h3 * w3;
    UmOperand A = um_kmajor(rows, true, false), Bo = um_kmajor(32, true, false);
    const uint32_t a_bytes = A.part_bytes * 2;
    n->l_conv3.cta0 = (int)pl.ctas.size(); n->l_conv3.njt = 32;
    n->l_conv3.stage_bytes = std::max<uint32_t>(a_bytes + Bo.part_bytes * 2, A.part_bytes + 16384);
    n->l_conv3.stages = std::min<int>(kStagesMax, (int)((226 * 1024 - 1024 - kCtlBytes) / n->l_conv3.stage_bytes));
    for (int p = 0; p < d.npass; ++p) {
      const int blob = d.pass_target[p] ? 1 : 0;
      for (int half = 0; half < 2; ++half) {
        UmProblem pr;
        memset(&pr, 0, sizeof(pr));
        pr.A = A; pr.B = Bo; pr.ksteps = 4; pr.run_stages = 1; pr.red_per_stage = 32; pr.epi = UM_EPI_ROWS;
        pr.MI = rows; pr.NJ = 32;
        pr.out_hi = n->act_hi[2] + 32 * half; pr.out_lo = n->act_lo[2] + 32 * half; pr.out_f32 = n->act_f32[2] + 32 * half;
        pr.out_ld = 64;
        pr.bias = (blob ? d.target : d.online) + d.off_conv_b[2] + 32 * half; pr.relu = 1;
        pr.pw = 1 << 20; pr.rs_outer = 0; pr.rs_inner = 1;
        const int prob = (int)pl.probs.size();
        pl.probs.push_back(pr);
        for (int t = 0; t < B / G; ++t) {
          const int img = p * B + t * G;
          UmCta c;
          memset(&c, 0, sizeof(c));
          c.prob = (uint32_t)prob; c.op0 = (uint32_t)pl.ops.size(); c.nstages = 18; c.ops_per_stage = 4;
          c.tx_bytes = (uint32_t)(2 * rows * 128 + 2 * 32 * 128);
          c.row_base = img * h3 * w3; c.ph_valid = 1; c.pw_valid = rows;
          for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw)
              for (int ch = 0; ch < 2; ++ch) {
                const int s = (kh * 3 + kw) * 2 + ch;
                for (int part = 0; part < 2; ++part) push_op(pl, m_a[part], part * A.part_bytes, 32 * ch, kw, kh, img, 0);
                for (int part = 0; part < 2; ++part) push_op(pl, m_wf[blob][2][part], a_bytes + part * Bo.part_bytes, 32 * s, 32 * half, 0, 0, 0);
              }
          pl.ctas.push_back(c);
        }
      }
    }
    n->l_conv3.nctas = (int)pl.ctas.size() - n->l_conv3.cta0;
  }

  // =========================================================================== conv3 input gradient
  // dact2[b,y,x,c] = [act2 > 0] * sum_{kh,kw,n} dact3[b, y-kh, x-kw, n] * W3[kh,kw,c,n]; halo zero-filled by the TMA unit
  {
    const int nb = 2;                                     // row bands per image
    const int hb = (h2 + nb - 1) / nb;
    int m_a[2], m_w[2];
    for (int part = 0; part < 2; ++part) {
      uint64_t dims[4] = {64, (uint64_t)w3, (uint64_t)h3, (uint64_t)B};
      uint64_t strides[3] = {256, (uint64_t)w3 * 256, (uint64_t)h3 * w3 * 256};
      uint32_t box[4] = {32, (uint32_t)w2, (uint32_t)hb, 1};
      m_a[part] = pl.add_map(part ? n->dact_lo[2] : n->dact_hi[2], 4, dims, strides, box);
      uint64_t wd[2] = {576, 64}, ws[1] = {576 * 4};
      uint32_t wb[2] = {32, 32};
      m_w[part] = pl.add_map(part ? n->wd3_lo : n->wd3_hi, 2, wd, ws, wb);
      if (m_a[part] < 0 || m_w[part] < 0) return DZ_EINVAL;
    }
    const int rows = hb * w2;
    UmOperand A = um_kmajor(rows, true, false), Bo = um_kmajor(32, true, false);
    const uint32_t a_bytes = A.part_bytes * 2;
    n->l_dconv3.cta0 = (int)pl.ctas.size(); n->l_dconv3.njt = 32;
    n->l_dconv3.stage_bytes = std::max<uint32_t>(a_bytes + Bo.part_bytes * 2, A.part_bytes + 16384);
    n->l_dconv3.stages = std::min<int>(kStagesMax, (int)((226 * 1024 - 1024 - kCtlBytes) / n->l_dconv3.stage_bytes));
    for (int half = 0; half < 2; ++half) {
      UmProblem pr;
      memset(&pr, 0, sizeof(pr));
      pr.A = A; pr.B = Bo; pr.ksteps = 4; pr.run_stages = 2; pr.red_per_stage = 32; pr.epi = UM_EPI_ROWS;
      pr.MI = rows; pr.NJ = 32;
      pr.out_hi = n->dact_hi[1] + 32 * half; pr.out_lo = n->dact_lo[1] + 32 * half; pr.out_f32 = n->dact_f32[1] + 32 * half;
      pr.out_ld = 64;
      pr.mask = n->act_hi[1] + 32 * half;                 // pass 0 is the first B images
      pr.pw = 1 << 20; pr.rs_outer = 0; pr.rs_inner = 1;
      const int prob = (int)pl.probs.size();
      pl.probs.push_back(pr);
      for (int b = 0; b < B; ++b)
        for (int band = 0; band < nb; ++band) {
          const int y0 = band * hb, y1 = std::min(h2, y0 + hb);
          if (y1 <= y0) continue;
          UmCta c;
          memset(&c, 0, sizeof(c));
          c.prob = (uint32_t)prob; c.op0 = (uint32_t)pl.ops.size(); c.nstages = 18; c.ops_per_stage = 4;
          c.tx_bytes = (uint32_t)(2 * rows * 128 + 2 * 32 * 128);
          c.row_base = (b * h2 + y0) * w2; c.ph_valid = 1; c.pw_valid = (y1 - y0) * w2;
          for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw)
              for (int nh = 0; nh < 2; ++nh) {
                const int s = (kh * 3 + kw) * 2 + nh;
                for (int part = 0; part < 2; ++part) push_op(pl, m_a[part], part * A.part_bytes, 32 * nh, -kw, y0 - kh, b, 0);
                for (int part = 0; part < 2; ++part) push_op(pl, m_w[part], a_bytes + part * Bo.part_bytes, 32 * s, 32 * half, 0, 0, 0);
              }
          pl.ctas.push_back(c);
        }
    }
    n->l_dconv3.nctas = (int)pl.ctas.size() - n->l_dconv3.cta0;
  }

  // =========================================================================== conv2 input gradient (4 parity classes)
  // dact1[b, 2i+py, 2j+px, c] = [act1 > 0] * sum_{ay,ax,n} dact2[b, i-ay, j-ax, n] * W2[py+2ay, px+2ax, c, n]
  {
    int m_a[2], m_w[2];
    for (int part = 0; part < 2; ++part) {
      uint64_t dims[4] = {64, (uint64_t)w2, (uint64_t)h2, (uint64_t)B};
      uint64_t strides[3] = {256, (uint64_t)w2 * 256, (uint64_t)h2 * w2 * 256};
      uint32_t box[4] = {32, (uint32_t)w1 / 2, (uint32_t)h1 / 2, 1};
      m_a[part] = pl.add_map(part ? n->dact_lo[1] : n->dact_hi[1], 4, dims, strides, box);
      uint64_t wd[2] = {256, 128}, ws[1] = {256 * 4};
      uint32_t wb[2] = {32, 32};
      m_w[part] = pl.add_map(part ? n->wd2_lo : n->wd2_hi, 2, wd, ws, wb);
      if (m_a[part] < 0 || m_w[part] < 0) return DZ_EINVAL;
    }
    const int rows = (h1 / 2) * (w1 / 2);
    UmOperand A = um_kmajor(rows, true, false), Bo = um_kmajor(32, true, false);
    const uint32_t a_bytes = A.part_bytes * 2;
    n->l_dconv2.cta0 = (int)pl.ctas.size(); n->l_dconv2.njt = 32;
    n->l_dconv2.stage_bytes = std::max<uint32_t>(a_bytes + Bo.part_bytes * 2, A.part_bytes + 16384);
    n->l_dconv2.stages = std::min<int>(kStagesMax, (int)((226 * 1024 - 1024 - kCtlBytes) / n->l_dconv2.stage_bytes));
    UmProblem pr;
    memset(&pr, 0, sizeof(pr));
    pr.A = A; pr.B = Bo; pr.ksteps = 4; pr.run_stages = 2; pr.red_per_stage = 32; pr.epi = UM_EPI_ROWS;
    pr.MI = rows; pr.NJ = 32;
    pr.out_hi = n->dact_hi[0]; pr.out_lo = n->dact_lo[0]; pr.out_f32 = n->dact_f32[0]; pr.out_ld = 32;
    pr.mask = n->act_hi[0];
    pr.pw = w1 / 2; pr.rs_outer = 2 * w1; pr.rs_inner = 2;
    const int prob = (int)pl.probs.size();
    pl.probs.push_back(pr);
    for (int b = 0; b < B; ++b)
      for (int cls = 0; cls < 4; ++cls) {
        const int py = cls >> 1, pxx = cls & 1;
        UmCta c;
        memset(&c, 0, sizeof(c));
        c.prob = (uint32_t)prob; c.op0 = (uint32_t)pl.ops.size(); c.nstages = 8; c.ops_per_stage = 4;
        c.tx_bytes = (uint32_t)(2 * rows * 128 + 2 * 32 * 128);
        c.row_base = b * h1 * w1 + py * w1 + pxx; c.ph_valid = h1 / 2; c.pw_valid = w1 / 2;
        for (int ay = 0; ay < 2; ++ay)
          for (int ax = 0; ax < 2; ++ax)
            for (int nh = 0; nh < 2; ++nh) {
              const int s = (ay * 2 + ax) * 2 + nh;
              for (int part = 0; part < 2; ++part) push_op(pl, m_a[part], part * A.part_bytes, 32 * nh, -ax, -ay, b, 0);
              for (int part = 0; part < 2; ++part) push_op(pl, m_w[part], a_bytes + part * Bo.part_bytes, 32 * s, 32 * cls, 0, 0, 0);
            }
        pl.ctas.push_back(c);
      }
    n->l_dconv2.nctas = (int)pl.ctas.size() - n->l_dconv2.cta0;
  }

  for (int part = 0; part < 2; ++part) {   // conv1 weight gradient: G operand
    uint64_t dims[2] = {32, (uint64_t)B * h1 * w1}, strides[1] = {128};
    uint32_t box[2] = {32, 128};
    n->map_g1[part] = pl.add_map(part ? n->dact_lo[0] : n->dact_hi[0], 2, dims, strides, box, true);
    if (n->map_g1[part] < 0) return DZ_EINVAL;
  }
  // =========================================================================== conv3 / conv2 weight gradients
  // dW[k][n] = sum_m A[m][k] G[m][n]: both operands are read through MN-major (transposing) descriptors straight from the
  // NHWC tensors — the same im2col boxes as the forward pass, with the reduction running over (output row, 8 images).
  {
    const int groups = (B + 7) / 8;
    // ---- conv3: A = act2 patches (pass 0), G = dact3
    {
      int m_a[2], m_g[2];
      for (int part = 0; part < 2; ++part) {
        uint64_t dims[4] = {64, (uint64_t)w2, (uint64_t)h2, (uint64_t)PB};
        uint64_t strides[3] = {256, (uint64_t)w2 * 256, (uint64_t)h2 * w2 * 256};
        uint32_t box[4] = {32, (uint32_t)w3, 1, 8};
        m_a[part] = pl.add_map(part ? n->act_lo[1] : n->act_hi[1], 4, dims, strides, box, true);
        uint64_t gd[4] = {64, (uint64_t)w3, (uint64_t)h3, (uint64_t)B};
        uint64_t gs[3] = {256, (uint64_t)w3 * 256, (uint64_t)h3 * w3 * 256};
        m_g[part] = pl.add_map(part ? n->dact_lo[2] : n->dact_hi[2], 4, gd, gs, box, true);
        if (m_a[part] < 0 || m_g[part] < 0) return DZ_EINVAL;
      }
      const int rows = w3 * 8;                                  // reduction rows per stage (multiple of 8)
      UmOperand A = um_mnmajor(128, rows, false), Bo = um_mnmajor(64, rows, false);
      const uint32_t a_bytes = A.part_bytes * 2;
      n->l_wconv3.cta0 = (int)pl.ctas.size(); n->l_wconv3.njt = 64;
      n->l_wconv3.stage_bytes = (a_bytes + Bo.part_bytes * 2 + 1023) / 1024 * 1024;
      n->l_wconv3.stages = std::max(1, std::min<int>(kStagesMax, (int)((226 * 1024 - 1024 - kCtlBytes) / n->l_wconv3.stage_bytes)));
      UmProblem pr;
      memset(&pr, 0, sizeof(pr));
      pr.A = A; pr.B = Bo; pr.ksteps = (uint32_t)(rows / 8); pr.run_stages = 1; pr.red_per_stage = (uint32_t)rows;
      pr.epi = UM_EPI_PARTIAL; pr.MI = 576; pr.NJ = 64;
      pr.C = n->wg3_part; pr.sc_i = 64; pr.sc_j = 1; pr.split_stride = 576 * 64;
      const int prob = (int)pl.probs.size();
      pl.probs.push_back(pr);
      const int S = n->wg3_splits, per = (h3 + S - 1) / S;
      for (int kt = 0; kt < 5; ++kt)                            // 18 slabs of 32 reduction... k values: 4,4,4,4,2
        for (int sp = 0; sp < S; ++sp) {
          const int oy0 = sp * per, oy1 = std::min(h3, oy0 + per);
          if (oy1 <= oy0) continue;
          UmCta c;
          memset(&c, 0, sizeof(c));
          const int nslab = std::min(4, 18 - kt * 4);
          c.prob = (uint32_t)prob; c.op0 = (uint32_t)pl.ops.size(); c.nstages = (uint32_t)((oy1 - oy0) * groups);
          c.ops_per_stage = (uint32_t)(2 * nslab + 4);
          c.tx_bytes = (uint32_t)((2 * nslab + 4) * rows * 128);
          c.i0 = kt * 128; c.split = sp;
          for (int oy = oy0; oy < oy1; ++oy)
            for (int gidx = 0; gidx < groups; ++gidx) {
              for (int part = 0; part < 2; ++part)
                for (int sl = 0; sl < nslab; ++sl) {
                  const int slab = kt * 4 + sl, tap = slab >> 1, ch = slab & 1, kh = tap / 3, kw = tap % 3;
                  push_op(pl, m_a[part], part * A.part_bytes + sl * A.lbo, 32 * ch, kw, oy + kh, gidx * 8, 0);
                }
              for (int part = 0; part < 2; ++part)
                for (int nh = 0; nh < 2; ++nh) push_op(pl, m_g[part], a_bytes + part * Bo.part_bytes + nh * Bo.lbo, 32 * nh, 0, oy, gidx * 8, 0);
            }
          pl.ctas.push_back(c);
        }
      n->l_wconv3.nctas = (int)pl.ctas.size() - n->l_wconv3.cta0;
    }
    // ---- conv2: A = act1 patches through the stride-2 parity view (pass 0), G = dact2; two 32-column halves
    {
      int m_a[2], m_g[2];
      for (int part = 0; part < 2; ++part) {
        uint64_t dims[5] = {64, (uint64_t)w1 / 2, 2, (uint64_t)h1 / 2, (uint64_t)PB};
        uint64_t strides[4] = {256, (uint64_t)w1 * 128, (uint64_t)2 * w1 * 128, (uint64_t)h1 * w1 * 128};
        uint32_t box[5] = {32, (uint32_t)w2, 1, 1, 8};
        m_a[part] = pl.add_map(part ? n->act_lo[0] : n->act_hi[0], 5, dims, strides, box, true);
        uint64_t gd[4] = {64, (uint64_t)w2, (uint64_t)h2, (uint64_t)B};
        uint64_t gs[3] = {256, (uint64_t)w2 * 256, (uint64_t)h2 * w2 * 256};
        uint32_t gbox[4] = {32, (uint32_t)w2, 1, 8};
        m_g[part] = pl.add_map(part ? n->dact_lo[1] : n->dact_hi[1], 4, gd, gs, gbox, true);
        if (m_a[part] < 0 || m_g[part] < 0) return DZ_EINVAL;
      }
      const int rows = w2 * 8;
      UmOperand A = um_mnmajor(128, rows, false), Bo = um_mnmajor(32, rows, false);
      const uint32_t a_bytes = A.part_bytes * 2;
      n->l_wconv2.cta0 = (int)pl.ctas.size(); n->l_wconv2.njt = 32;
      n->l_wconv2.stage_bytes = (a_bytes + Bo.part_bytes * 2 + 1023) / 1024 * 1024;
      n->l_wconv2.stages = std::max(1, std::min<int>(kStagesMax, (int)((226 * 1024 - 1024 - kCtlBytes) / n->l_wconv2.stage_bytes)));
      const int S = n->wg2_splits, per = (h2 + S - 1) / S;
      for (int half = 0; half < 2; ++half) {
        UmProblem pr;
        memset(&pr, 0, sizeof(pr));
        pr.A = A; pr.B = Bo; pr.ksteps = (uint32_t)(rows / 8); pr.run_stages = 1; pr.red_per_stage = (uint32_t)rows;
        pr.epi = UM_EPI_PARTIAL; pr.MI = 512; pr.NJ = 32;
        pr.C = n->wg2_part + 32 * half; pr.sc_i = 64; pr.sc_j = 1; pr.split_stride = 512 * 64;
        const int prob = (int)pl.probs.size();
        pl.probs.push_back(pr);
        for (int kt = 0; kt < 4; ++kt)
          for (int sp = 0; sp < S; ++sp) {
            const int oy0 = sp * per, oy1 = std::min(h2, oy0 + per);
            if (oy1 <= oy0) continue;
            UmCta c;
            memset(&c, 0, sizeof(c));
            c.prob = (uint32_t)prob; c.op0 = (uint32_t)pl.ops.size(); c.nstages = (uint32_t)((oy1 - oy0) * groups);
            c.ops_per_stage = 10;
            c.tx_bytes = (uint32_t)(10 * rows * 128);
            c.i0 = kt * 128; c.split = sp;
            for (int oy = oy0; oy < oy1; ++oy)
              for (int gidx = 0; gidx < groups; ++gidx) {
                for (int part = 0; part < 2; ++part)
                  for (int sl = 0; sl < 4; ++sl) {
                    const int tap = kt * 4 + sl, kh = tap >> 2, kw = tap & 3;   // one slab = one (kh, kw) tap x 32 channels
                    push_op(pl, m_a[part], part * A.part_bytes + sl * A.lbo, 32 * (kw & 1), kw >> 1, kh & 1, oy + (kh >> 1), gidx * 8);
                  }
                for (int part = 0; part < 2; ++part) push_op(pl, m_g[part], a_bytes + part * Bo.part_bytes, 32 * half, 0, oy, gidx * 8, 0);
              }
            pl.ctas.push_back(c);
          }
      }
      n->l_wconv2.nctas = (int)pl.ctas.size() - n->l_wconv2.cta0;
    }
  }

  // =========================================================================== fc1 / noisy1 forward and input gradient
  n->l_fc.nctas = 0; n->l_fcd.nctas = 0;
  if (d.use_fc) {
    const int njt = B <= 32 ? 32 : 64;
    n->njt_fc = njt;
    const int q = d.noisy ? 2 : 1;
    int m_x[2], m_g[2];
    for (int part = 0; part < 2; ++part) {
      uint64_t dims[2] = {(uint64_t)feat, (uint64_t)PB}, strides[1] = {(uint64_t)feat * 4};
      uint32_t box[2] = {32, (uint32_t)njt};
      m_x[part] = pl.add_map(part ? n->act_lo[2] : n->act_hi[2], 2, dims, strides, box);
      uint64_t gd[2] = {512, (uint64_t)d.nstream * B}, gs[1] = {2048};
      m_g[part] = pl.add_map(part ? n->dh1_lo : n->dh1_hi, 2, gd, gs, box);
      if (m_x[part] < 0 || m_g[part] < 0) return DZ_EINVAL;
    }
    // weight maps: [blob][stream][sigma] x {forward box (32 n, 32 k), gradient box (32 n, 128 k)}
    int m_wf_fc[2][2][2], m_wd_fc[2][2];
    bool wf3d = true;
    for (int blob = 0; blob < 2; ++blob)
      for (int s = 0; s < d.nstream; ++s)
        for (int sg = 0; sg < q; ++sg) {
          const float* w = (blob ? d.target : d.online) + (sg ? d.off_fc_sw[s] : d.off_fc_w[s]);
          uint64_t dims[2] = {512, (uint64_t)feat}, strides[1] = {2048};
          // MN-major A operand (rows of W are the reduction): W[k][n] viewed as (n % 32, k, n / 32) so that ONE box of
          // (32, 32, 4) lands as the four [32 k][32 n] slabs of a 128-column tile, slab-major — one TMA op per 16 KB tile
          uint64_t dims3[3] = {32, (uint64_t)feat, 16}, strides3[2] = {2048, 128};
          uint32_t box3[3] = {32, 32, 4};
          if (wf3d) {
            m_wf_fc[blob][s][sg] = pl.add_map(w, 3, dims3, strides3, box3, true);
            if (m_wf_fc[blob][s][sg] < 0) wf3d = false;      // driver refused the permuted view: one op per slab instead
          }
          if (!wf3d) {
            if (blob || s || sg) return fail(DZ_EINVAL, "fc weight tensor maps: inconsistent encodings");
            uint32_t box2[2] = {32, 32};
            m_wf_fc[blob][s][sg] = pl.add_map(w, 2, dims, strides, box2, true);
          }
          if (m_wf_fc[blob][s][sg] < 0) return DZ_EINVAL;
          if (blob == 0) {
            uint32_t boxd[2] = {32, 128};
            m_wd_fc[s][sg] = pl.add_map(w, 2, dims, strides, boxd);
            if (m_wd_fc[s][sg] < 0) return DZ_EINVAL;
          }
        }
    // ---- forward: D[n, m] = sum_k W[k][n] x[m][k];  noisy: W = Wmu + Wsigma * (eps_in[k] * eps_out[n]) formed by the converters
    {
      UmOperand Bo = um_kmajor(njt, true, false);
      const int nk = feat / 32, S = n->fc_splits, per = (nk + S - 1) / S;
      n->l_fc.cta0 = (int)pl.ctas.size(); n->l_fc.njt = njt; n->l_fc.convert = true;
      n->l_fc.stage_bytes = 2 * 16384 + 2 * Bo.part_bytes;
      n->l_fc.stages = std::min<int>(kStagesMax, (int)((226 * 1024 - 1024 - kCtlBytes) / n->l_fc.stage_bytes));
      for (int p = 0; p < d.npass; ++p) {
        const int blob = d.pass_target[p] ? 1 : 0;
        for (int s = 0; s < d.nstream; ++s) {
          const int qi = p * d.nstream + s;
          UmProblem pr;
          memset(&pr, 0, sizeof(pr));
          pr.A = um_mnmajor(128, 32, true, nullptr); pr.B = Bo; pr.ksteps = 4; pr.run_stages = 1; pr.red_per_stage = 32;
          if (d.noisy) pr.A.convert = 2;
          pr.epi = UM_EPI_PARTIAL; pr.MI = 512; pr.NJ = B;
          pr.C = n->fc_part + (int64_t)qi * S * B * 512; pr.sc_i = 1; pr.sc_j = 512; pr.split_stride = (long long)B * 512;
          const int prob = (int)pl.probs.size();
          pl.probs.push_back(pr);
          if (d.noisy) {
            n->patches.push_back({prob, 0, (int64_t)d.noise_apply[p] * d.noise_stride + d.noise_off_in[s]});
            n->patches.push_back({prob, 2, (int64_t)d.noise_apply[p] * d.noise_stride + d.noise_off_out[s]});
          }
          // the four 128-column tiles of one k range sit in neighbouring CTAs: together they stream whole 2 KB rows of W
          for (int sp = 0; sp < S; ++sp)
            for (int nt = 0; nt < 4; ++nt) {
              const int k0 = sp * per, k1 = std::min(nk, k0 + per);
              if (k1 <= k0) continue;
              UmCta c;
              memset(&c, 0, sizeof(c));
              c.prob = (uint32_t)prob; c.op0 = (uint32_t)pl.ops.size(); c.nstages = (uint32_t)(k1 - k0);
              c.ops_per_stage = (wf3d ? 1 : 4) * (d.noisy ? 2 : 1) + 2;
              c.tx_bytes = (uint32_t)((d.noisy ? 32768 : 16384) + 2 * njt * 128);
              c.r0 = 32 * k0; c.i0 = nt * 128; c.split = sp;
              for (int ks = k0; ks < k1; ++ks) {
                for (int sg = 0; sg < (d.noisy ? 2 : 1); ++sg) {
                  if (wf3d) push_op(pl, m_wf_fc[blob][s][sg], sg * 16384, 0, 32 * ks, nt * 4, 0, 0);
                  else for (int sl = 0; sl < 4; ++sl) push_op(pl, m_wf_fc[blob][s][sg], sg * 16384 + sl * 4096, nt * 128 + 32 * sl, 32 * ks, 0, 0, 0);
                }
                for (int part = 0; part < 2; ++part) push_op(pl, m_x[part], 32768 + part * Bo.part_bytes, 32 * ks, p * B, 0, 0, 0);
              }
              pl.ctas.push_back(c);
            }
        }
      }
      n->l_fc.nctas = (int)pl.ctas.size() - n->l_fc.cta0;
    }
    // ---- input gradient: D[k, m] = sum_n W[k][n] g[m][n];  noisy: W = Wmu + Wsigma * (eps_in[k] * eps_out[n]) in the converters
    {
      UmOperand Bo = um_kmajor(njt, true, false);
      const int S = n->fcd_splits, per = (16 + S - 1) / S, ktiles = (feat + 127) / 128;
      n->l_fcd.cta0 = (int)pl.ctas.size(); n->l_fcd.njt = njt; n->l_fcd.convert = true;
      n->l_fcd.stage_bytes = 2 * 16384 + 2 * Bo.part_bytes;
      n->l_fcd.stages = std::min<int>(kStagesMax, (int)((226 * 1024 - 1024 - kCtlBytes) / n->l_fcd.stage_bytes));
      for (int s = 0; s < d.nstream; ++s) {
        UmProblem pr;
        memset(&pr, 0, sizeof(pr));
        pr.A = um_kmajor(128, true, true, nullptr); pr.B = Bo; pr.ksteps = 4; pr.run_stages = 2; pr.red_per_stage = 32;
        if (d.noisy) pr.A.convert = 2;
        pr.epi = UM_EPI_PARTIAL; pr.MI = feat; pr.NJ = B;
        pr.C = n->fcd_part + (int64_t)s * S * B * feat; pr.sc_i = 1; pr.sc_j = feat; pr.split_stride = (long long)B * feat;
        const int prob = (int)pl.probs.size();
        pl.probs.push_back(pr);
        if (d.noisy) {
          n->patches.push_back({prob, 0, (int64_t)d.noise_apply[0] * d.noise_stride + d.noise_off_out[s]});
          n->patches.push_back({prob, 2, (int64_t)d.noise_apply[0] * d.noise_stride + d.noise_off_in[s]});
        }
        for (int kt = 0; kt < ktiles; ++kt)
          for (int sp = 0; sp < S; ++sp) {
            const int n0 = sp * per, n1 = std::min(16, n0 + per);
            if (n1 <= n0) continue;
            UmCta c;
            memset(&c, 0, sizeof(c));
            c.prob = (uint32_t)prob; c.op0 = (uint32_t)pl.ops.size(); c.nstages = (uint32_t)(n1 - n0);
            c.ops_per_stage = d.noisy ? 4 : 3;
            c.tx_bytes = (uint32_t)((d.noisy ? 32768 : 16384) + 2 * njt * 128);
            c.r0 = 32 * n0; c.i0 = kt * 128; c.split = sp;
            for (int ns = n0; ns < n1; ++ns) {
              push_op(pl, m_wd_fc[s][0], 0, 32 * ns, kt * 128, 0, 0, 0);
              if (d.noisy) push_op(pl, m_wd_fc[s][1], 16384, 32 * ns, kt * 128, 0, 0, 0);
              for (int part = 0; part < 2; ++part) push_op(pl, m_g[part], 32768 + part * Bo.part_bytes, 32 * ns, s * B, 0, 0, 0);
            }
            pl.ctas.push_back(c);
          }
      }
      n->l_fcd.nctas = (int)pl.ctas.size() - n->l_fcd.cta0;
    }
  }
  return DZ_OK;
}

int apply_noise(UmNet* n, const float* noise, void* stream) {
  if (n->patches.empty() || noise == n->noise_cached) return DZ_OK;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing((cudaStream_t)stream, &cs);
  if (cs != cudaStreamCaptureStatusNone) return fail(DZ_EINVAL, "the noise buffer must not move between graph captures (keep one buffer per learner)");
  for (const auto& p : n->patches) {
    UmProblem& pr = n->plan.probs[p.prob];
    if (p.field == 0) pr.A.scale_r = noise + p.off; else if (p.field == 2) pr.A.scale_i = noise + p.off; else pr.scale_i = noise + p.off;
  }
  DZ_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  DZ_CUDA_OK(cudaMemcpy(n->plan.d_probs, n->plan.probs.data(), n->plan.probs.size() * sizeof(UmProblem), cudaMemcpyHostToDevice));
  n->noise_cached = noise;
  return DZ_OK;
}

}  // namespace

int64_t um_net_workspace_bytes(const UmNetDesc& d) {
  UmNet tmp;
  tmp.d = d;
  return carve_net(&tmp, nullptr);
}

int um_net_create(const UmNetDesc& d, char* base, UmNet** out) {
  if (!um_net_supported(d)) return fail(DZ_EINVAL, "geometry not supported by the tcgen05 path");
  UmNet* n = new UmNet();
  n->d = d;
  carve_net(n, base);
  // conv1 geometry
  const int px = n->h1 * n->w1, m_pass = d.B * px;
  n->conv1_tiles_per_pass = (m_pass + 127) / 128;
  int worst = 0;
  for (int m0 = 0; m0 < m_pass; m0 += 128) {
    const int m1 = std::min(m0 + 128, m_pass);
    int tot = 0;
    for (int b = m0 / px; b <= (m1 - 1) / px; ++b) {
      const int plo = std::max(m0, b * px) - b * px, phi = std::min(m1, (b + 1) * px) - b * px;
      tot += (4 * ((phi - 1) / n->w1 - plo / n->w1) + 8) * d.W * 4;
    }
    worst = std::max(worst, tot);
  }
  n->conv1_stag_bytes = (worst + 127) / 128 * 128;
  cudaMemset(n->wg_ticket, 0, 64 * 4);
  // gradient buffers start as zeros (hi/lo pairs of layers whose producer has not run yet are never NaN)
  for (int L = 0; L < 3; ++L) {
    const int64_t cnt[3] = {(int64_t)d.B * n->h1 * n->w1 * 32, (int64_t)d.B * n->h2 * n->w2 * 64, (int64_t)d.B * n->feat};
    cudaMemset(n->dact_hi[L], 0, cnt[L] * 4); cudaMemset(n->dact_lo[L], 0, cnt[L] * 4); cudaMemset(n->dact_f32[L], 0, cnt[L] * 4);
  }
  int rc = build_plan(n);
  UmLaunch* all[8] = {&n->l_conv2, &n->l_conv3, &n->l_dconv3, &n->l_dconv2, &n->l_fc, &n->l_fcd, &n->l_wconv3, &n->l_wconv2};
  for (int i = 0; i < 8 && rc == DZ_OK; ++i)
    if (all[i]->nctas > 0) rc = n->plan.localize_maps(*all[i]);
  if (rc == DZ_OK) rc = n->plan.upload();
  if (rc == DZ_OK) rc = UmPlan::configure();
  if (rc != DZ_OK) { um_net_destroy(n); return rc; }
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(conv1_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(conv1_wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      um_net_destroy(n);
      return fail(DZ_ECUDA, "conv1_umma_kernel shared memory attribute");
    }
    configured = true;
  }
  *out = n;
  return DZ_OK;
}

void um_net_trace(UmNet* n, const char* tag, long long* d_trace) { n->trace_tag = tag ? tag : ""; n->trace_ptr = d_trace; }

void um_net_destroy(UmNet* n) {
  if (!n) return;
  n->plan.release();
  delete n;
}

// Layers 1 and 2 have no consumer of the exact fp32 activation on this path (the next layer reads the hi/lo pair, the
// ReLU masks read the sign of hi): their "fp32 view" is the tf32 hi image — same activation pattern, values rounded to tf32.
float* um_act_f32(UmNet* n, int layer, int pass) {
  const int64_t per[3] = {(int64_t)n->d.B * n->h1 * n->w1 * 32, (int64_t)n->d.B * n->h2 * n->w2 * 64, (int64_t)n->d.B * n->feat};
  return (layer < 3 ? n->act_hi[layer - 1] : n->act_f32[layer - 1]) + per[layer - 1] * pass;
}
float* um_dact_f32(UmNet* n, int layer) { return n->dact_f32[layer - 1]; }
float* um_h1_f32(UmNet* n, int pass, int stream) { return n->h1_buf + ((int64_t)pass * n->d.nstream + stream) * n->d.B * 512; }
float* um_dh1_f32(UmNet* n, int stream) { return n->dh1_f32 + (int64_t)stream * n->d.B * 512; }
float* um_dh1_hi(UmNet* n, int stream) { return n->dh1_hi + (int64_t)stream * n->d.B * 512; }
float* um_dh1_lo(UmNet* n, int stream) { return n->dh1_lo + (int64_t)stream * n->d.B * 512; }

int um_pack_weights(UmNet* n, void* stream) {
  PackArgs a;
  memset(&a, 0, sizeof(a));
  a.blob[0] = n->d.online; a.blob[1] = n->d.target;
  for (int L = 0; L < 3; ++L) a.off_w[L] = n->d.off_conv_w[L];
  for (int b = 0; b < 2; ++b)
    for (int L = 0; L < 3; ++L) { a.wf_hi[b][L] = n->wf_hi[b][L]; a.wf_lo[b][L] = n->wf_lo[b][L]; }
  a.wd3_hi = n->wd3_hi; a.wd3_lo = n->wd3_lo; a.wd2_hi = n->wd2_hi; a.wd2_lo = n->wd2_lo;
  const int total = 2 * 77824 + 64 * 576 + 128 * 256;
  DZ_LAUNCH_NAMED("conv_pack", um_pack_conv_kernel, (unsigned)ceil_div(total, 256), 256, 0, stream, a);
  return DZ_OK;
}

int um_prefetch_fc(UmNet* n, void* stream) {
  const UmNetDesc& d = n->d;
  if (!d.use_fc) return DZ_OK;
  PrefetchArgs a;
  memset(&a, 0, sizeof(a));
  const long long lines = ((long long)n->feat * 512 * 4 + 127) / 128;
  for (int blob = 0; blob < 2; ++blob)
    for (int s = 0; s < d.nstream; ++s)
      for (int sg = 0; sg < (d.noisy ? 2 : 1); ++sg) {
        a.ptr[a.n] = (blob ? d.target : d.online) + (sg ? d.off_fc_sw[s] : d.off_fc_w[s]);
        a.lines[a.n++] = lines;
      }
  DZ_LAUNCH_NAMED("fc_prefetch", um_prefetch_kernel, 148 * 4, 256, 0, stream, a);
  return DZ_OK;
}

int um_forward_torso(UmNet* n, const uint8_t* const* const* rows, void* stream) {
  const UmNetDesc& d = n->d;
  Conv1Args a;
  memset(&a, 0, sizeof(a));
  for (int p = 0; p < d.npass; ++p) {
    const int blob = d.pass_target[p] ? 1 : 0;
    a.rows[p] = rows[p];
    a.bias[p] = (blob ? d.target : d.online) + d.off_conv_b[0];
    a.blob[p] = blob;
  }
  for (int b = 0; b < 2; ++b)
    for (int part = 0; part < 2; ++part) a.wmap[b][part] = n->plan.maps[n->map_wf1[b][part]];
  a.out_hi = n->act_hi[0]; a.out_lo = n->act_lo[0]; a.out_f32 = n->act_f32[0];
  a.npass = d.npass; a.B = d.B; a.W = d.W; a.oh = n->h1; a.ow = n->w1; a.m_pass = d.B * n->h1 * n->w1;
  a.tiles_per_pass = n->conv1_tiles_per_pass; a.ntiles = a.tiles_per_pass * d.npass; a.stag_bytes = n->conv1_stag_bytes;
  a.trace = n->tr("conv1_fwd");
  const size_t smem = 2048 + kC1W + kC1A + 2 * (size_t)a.stag_bytes + kC1Epi;
  if (smem > 227 * 1024) return fail(DZ_EINVAL, "conv1 staging does not fit");
  const unsigned grid = (unsigned)std::min(148, a.ntiles);
  DZ_LAUNCH_NAMED("conv1_fwd", conv1_umma_kernel, grid, kThreadsU, smem, stream, a);
  DZ_TRY_RC(n->plan.launch("conv2_fwd", n->l_conv2, stream, n->tr("conv2_fwd")));
  DZ_TRY_RC(n->plan.launch("conv3_fwd", n->l_conv3, stream, n->tr("conv3_fwd")));
  return DZ_OK;
}

int um_forward_fc(UmNet* n, const float* noise, void* stream) {
  const UmNetDesc& d = n->d;
  if (!d.use_fc) return fail(DZ_EINVAL, "fc layers are not on the tcgen05 path for this agent");
  if (d.noisy) DZ_TRY_RC(apply_noise(n, noise, stream));
  DZ_TRY_RC(n->plan.launch(d.noisy ? "noisy1_fwd" : "fc1_fwd", n->l_fc, stream, n->tr("fc1_fwd")));
  FcFinishArgs a;
  memset(&a, 0, sizeof(a));
  a.part = n->fc_part; a.S = n->fc_splits; a.B = d.B; a.nstream = d.nstream; a.noisy = d.noisy; a.npass = d.npass; a.h1 = n->h1_buf;
  for (int p = 0; p < d.npass; ++p)
    for (int s = 0; s < d.nstream; ++s) {
      const float* blob = d.pass_target[p] ? d.target : d.online;
      a.bmu[p][s] = blob + d.off_fc_b[s];
      if (d.noisy) {
        a.bsig[p][s] = blob + d.off_fc_sb[s];
        a.eps_out[p][s] = noise + (int64_t)d.noise_apply[p] * d.noise_stride + d.noise_off_out[s];
      }
    }
  dim3 grid((unsigned)std::min<int64_t>(ceil_div((int64_t)d.B * 128, 256), 64), (unsigned)(d.npass * d.nstream));
  DZ_LAUNCH_NAMED("fc_finish", um_fc_finish_kernel, grid, 256, 0, stream, a);
  return DZ_OK;
}

int um_split_dh1(UmNet* n, void* stream) {
  return um_split(n->dh1_f32, n->dh1_hi, n->dh1_lo, (long long)n->d.nstream * n->d.B * 512, stream);
}

int um_backward_fc(UmNet* n, const float* noise, void* stream) {
  const UmNetDesc& d = n->d;
  if (!d.use_fc) return fail(DZ_EINVAL, "fc layers are not on the tcgen05 path for this agent");
  if (d.noisy) DZ_TRY_RC(apply_noise(n, noise, stream));
  DZ_TRY_RC(n->plan.launch(d.noisy ? "noisy1_dgrad" : "fc1_dgrad", n->l_fcd, stream, n->tr("fc1_dgrad")));
  const long long total = (long long)d.B * n->feat;
  DZ_LAUNCH_NAMED("fcd_finish", um_fcd_finish_kernel, (unsigned)std::min<long long>(ceil_div(total / 4, 256), 148 * 4), 256, 0, stream,
                  n->fcd_part, n->fcd_nsrc * n->fcd_splits, total, n->act_hi[2], n->dact_f32[2], n->dact_hi[2], n->dact_lo[2], total / 4);
  return DZ_OK;
}

int um_split_dact3(UmNet* n, void* stream) {
  return um_split(n->dact_f32[2], n->dact_hi[2], n->dact_lo[2], (long long)n->d.B * n->feat, stream);
}

int um_wgrad_conv1(UmNet* n, const uint8_t* const* rows0, void* stream) {
  const UmNetDesc& d = n->d;
  Conv1WgArgs a;
  memset(&a, 0, sizeof(a));
  a.rows = rows0; a.gmap[0] = n->plan.maps[n->map_g1[0]]; a.gmap[1] = n->plan.maps[n->map_g1[1]]; a.partial = n->wg1_part;
  a.B = d.B; a.W = d.W; a.oh = n->h1; a.ow = n->w1; a.m_pass = d.B * n->h1 * n->w1;
  a.ntiles = (a.m_pass + 127) / 128; a.stag_bytes = n->conv1_stag_bytes;
  const size_t smem = 2048 + kC1G + kC1A + 2 * (size_t)a.stag_bytes;
  if (smem > 227 * 1024) return fail(DZ_EINVAL, "conv1 wgrad staging does not fit");
  DZ_LAUNCH_NAMED("conv1_wgrad", conv1_wgrad_umma_kernel, (unsigned)n->wg1_ctas, kThreadsU, smem, stream, a);
  return DZ_OK;
}

int um_wgrad_conv3(UmNet* n, void* stream) { return n->plan.launch("conv3_wgrad", n->l_wconv3, stream, n->tr("conv3_wgrad")); }
int um_wgrad_conv2(UmNet* n, void* stream) { return n->plan.launch("conv2_wgrad", n->l_wconv2, stream, n->tr("conv2_wgrad")); }

// dW / db of conv3 and conv2 from the split partials (+ optionally conv1's FMA partials in the same launch).
namespace {
int wg_sum_blocks(int layer) { const int KN[3] = {256 * 32, 512 * 64, 576 * 64}; return (KN[layer - 1] / 4 + 31) / 32; }
int wg_slot_base(int layer) { int b = 0; for (int L = 3; L > layer; --L) b += wg_sum_blocks(L) + 1; return b; }
}  // namespace

int um_norm_slots(UmNet*) { return wg_slot_base(1) + wg_sum_blocks(1) + 1; }

// dW / db of one conv layer (1..3) from its split partials; norm_parts (optional): base of the split-norm slot array.
int um_wgrad_finish_layer(UmNet* n, int layer, float* dW, float* db, float* norm_parts, void* stream) {
  float* sc = n->wg_scratch;
  WgFinish f;
  memset(&f, 0, sizeof(f));
  if (layer == 3) f = WgFinish{n->wg3_part, n->wg3_splits, 576 * 64, 576 * 64, dW, n->dact_f32[2], n->d.B * n->h3 * n->w3, 64, db, sc, n->wg_ticket, 0, nullptr};
  else if (layer == 2) f = WgFinish{n->wg2_part, n->wg2_splits, 512 * 64, 512 * 64, dW, n->dact_f32[1], n->d.B * n->h2 * n->w2, 64, db, sc + 256 * 64, n->wg_ticket + 1, 0, nullptr};
  else f = WgFinish{n->wg1_part, n->wg1_ctas, 256 * 32, 256 * 32, dW, n->dact_f32[0], n->d.B * n->h1 * n->w1, 32, db, sc + 2 * 256 * 64, n->wg_ticket + 2, 0, nullptr};
  f.sum_blocks = wg_sum_blocks(layer);
  f.norm_part = norm_parts ? norm_parts + wg_slot_base(layer) : nullptr;
  const int chunks = (f.M + kWgChunkRows - 1) / kWgChunkRows;
  if (chunks > kWgMaxChunks) return fail(DZ_EINVAL, "bias-gradient reduction: too many row chunks");
  const char* tags[3] = {"wgrad_finish1", "wgrad_finish2", "wgrad_finish3"};
  DZ_LAUNCH_NAMED(tags[layer - 1], um_wgrad_finish_kernel, (unsigned)(f.sum_blocks + chunks), 256, 0, stream, f);
  return DZ_OK;
}

int um_backward_conv3(UmNet* n, void* stream) { return n->plan.launch("conv3_dgrad", n->l_dconv3, stream, n->tr("conv3_dgrad")); }
int um_backward_conv2(UmNet* n, void* stream) { return n->plan.launch("conv2_dgrad", n->l_dconv2, stream, n->tr("conv2_dgrad")); }

}  // namespace dz
