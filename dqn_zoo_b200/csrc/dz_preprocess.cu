// Atari frame preprocessing on the device (reference: dqn_zoo/processors.py:367-388 rgb2y + resize, :482-501 the
// observation branch of processors.atari()): for every environment stream
//     new = PillowBilinearResize( uint8( luma( max(frame_a, frame_b) ) ) )          (uint8 [out_h][out_w])
// pushed into that stream's frame stack uint8 [out_h][out_w][stack] (a deque: append while it is filling, shift
// left by one channel once it is full; processors.py:492-500).
//
// Byte/integer work, bound by the 2 x 100 KB of raw frame reads per stream: one CTA owns a band of output rows,
// stages the byte-wise max of the two raw frames for exactly the input rows that band needs (32-bit __vmaxu4
// loads), converts to luma in float64 with the reference's rounding order (see oracle/processors_oracle.py:rgb2y —
// products rounded separately, summed left to right, truncated), then runs Pillow's two fixed-point passes
// (22-bit coefficients, int32 accumulators, uint8 intermediate image) out of shared memory.
#include "dz_common.cuh"

namespace dz {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Pillow Resample.c PRECISION_BITS for 8-bit images
constexpr int kBandRows = 12;                // output rows per CTA

struct PreprocessArgs {
  const uint8_t* const* frame_a;
  const uint8_t* const* frame_b;
  dz_resample_axis h, v;
  uint8_t* const* stacks;
  const int32_t* counts;
  int stack;
  double wr, wg, wb;
  int max_rows;                              // input rows staged per band (shared-memory carve)
};

__device__ __forceinline__ uint8_t clip8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

__global__ void __launch_bounds__(256) atari_preprocess_kernel(const PreprocessArgs a) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int env = blockIdx.y;
  const int y0 = blockIdx.x * kBandRows, y1 = min(y0 + kBandRows, a.v.out_size);
  const int in_w = a.h.in_size, out_w = a.h.out_size;
  const int row_bytes = in_w * 3, row_words = row_bytes >> 2;
  // input rows this band touches: windows of consecutive output rows are monotone
  const int r0 = a.v.d_bounds[2 * y0];
  const int r1 = a.v.d_bounds[2 * (y1 - 1)] + a.v.d_bounds[2 * (y1 - 1) + 1];
  const int rows = r1 - r0;
  uint32_t* raw = reinterpret_cast<uint32_t*>(smem);                         // [max_rows][row_words]  pooled RGB
  uint8_t* gray = smem + (size_t)a.max_rows * row_bytes;                     // [max_rows][in_w]
  uint8_t* hpass = gray + (size_t)a.max_rows * in_w;                         // [max_rows][out_w]

  const uint8_t* fa = a.frame_a[env];
  const uint8_t* fb = a.frame_b[env];
  const uint32_t* wa = fa ? reinterpret_cast<const uint32_t*>(fa + (size_t)r0 * row_bytes) : nullptr;
  const uint32_t* wb = fb ? reinterpret_cast<const uint32_t*>(fb + (size_t)r0 * row_bytes) : nullptr;
  for (int i = threadIdx.x; i < rows * row_words; i += blockDim.x) {
    uint32_t x = wa ? wa[i] : 0u, y = wb ? wb[i] : 0u;
    raw[i] = __vmaxu4(x, y);                                                 // np.max over the pooled pair (:487)
  }
  __syncthreads();
  const uint8_t* rawb = reinterpret_cast<const uint8_t*>(raw);
  for (int i = threadIdx.x; i < rows * in_w; i += blockDim.x) {
    const uint8_t* px = rawb + (size_t)i * 3;
    double t = __dadd_rn(__dadd_rn(__dmul_rn((double)px[0], a.wr), __dmul_rn((double)px[1], a.wg)), __dmul_rn((double)px[2], a.wb));
    gray[i] = (uint8_t)(int)t;                                               // astype(np.uint8): truncation (:371)
  }
  __syncthreads();
  for (int i = threadIdx.x; i < rows * out_w; i += blockDim.x) {             // horizontal pass
    const int r = i / out_w, xx = i - r * out_w;
    const int xmin = a.h.d_bounds[2 * xx], cnt = a.h.d_bounds[2 * xx + 1];
    const int32_t* k = a.h.d_kk + (size_t)xx * a.h.ksize;
    const uint8_t* src = gray + (size_t)r * in_w + xmin;
    int acc = 1 << (kPrecisionBits - 1);
    for (int x = 0; x < cnt; ++x) acc += (int)src[x] * k[x];
    hpass[i] = clip8(acc >> kPrecisionBits);
  }
  __syncthreads();
  const int count = a.counts[env];
  uint8_t* stack = a.stacks[env];
  for (int i = threadIdx.x; i < (y1 - y0) * out_w; i += blockDim.x) {        // vertical pass + push into the stack
    const int yy = y0 + i / out_w, xx = i % out_w;
    const int ymin = a.v.d_bounds[2 * yy], cnt = a.v.d_bounds[2 * yy + 1];
    const int32_t* k = a.v.d_kk + (size_t)yy * a.v.ksize;
    int acc = 1 << (kPrecisionBits - 1);
    for (int y = 0; y < cnt; ++y) acc += (int)hpass[(size_t)(ymin - r0 + y) * out_w + xx] * k[y];
    const uint8_t v = clip8(acc >> kPrecisionBits);
    uint8_t* px = stack + ((size_t)yy * out_w + xx) * a.stack;
    if (count < a.stack) {
      px[count] = v;                                                         // still filling: trailing channels stay zero
    } else {
      for (int c = 0; c + 1 < a.stack; ++c) px[c] = px[c + 1];               // deque(maxlen): drop the oldest frame
      px[a.stack - 1] = v;
    }
  }
}

}  // namespace
}  // namespace dz

using namespace dz;

extern "C" int dz_atari_preprocess(const uint8_t* const* d_frame_a, const uint8_t* const* d_frame_b, int32_t n_env,
                                   const dz_resample_axis* horizontal, const dz_resample_axis* vertical,
                                   uint8_t* const* d_stacks, const int32_t* d_counts, int32_t stack, const double* luma3,
                                   int32_t max_band_rows, void* stream) {
  if (n_env <= 0) return DZ_OK;
  if (!d_frame_a || !d_frame_b || !horizontal || !vertical || !d_stacks || !d_counts || !luma3)
    return fail(DZ_EINVAL, "dz_atari_preprocess: null argument");
  if ((horizontal->in_size * 3) % 4) return fail(DZ_EINVAL, "dz_atari_preprocess: row bytes must be a multiple of 4");
  if (stack < 1) return fail(DZ_EINVAL, "dz_atari_preprocess: stack geometry");
  if (max_band_rows < 1 || max_band_rows > vertical->in_size) return fail(DZ_EINVAL, "dz_atari_preprocess: max_band_rows");
  PreprocessArgs a;
  a.frame_a = d_frame_a; a.frame_b = d_frame_b; a.h = *horizontal; a.v = *vertical; a.stacks = d_stacks;
  a.counts = d_counts; a.stack = stack; a.wr = luma3[0]; a.wg = luma3[1]; a.wb = luma3[2];
  a.max_rows = max_band_rows;
  const size_t smem = (size_t)max_band_rows * (horizontal->in_size * 3 + horizontal->in_size + horizontal->out_size);
  if (smem > 200 * 1024) return fail(DZ_EINVAL, "dz_atari_preprocess: band does not fit in shared memory");
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    DZ_CUDA_OK(cudaFuncSetAttribute(atari_preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid((unsigned)ceil_div(vertical->out_size, kBandRows), (unsigned)n_env);
  DZ_LAUNCH(atari_preprocess_kernel, grid, 256, smem, stream, a);
  return DZ_OK;
}

extern "C" int32_t dz_atari_preprocess_band_rows(void) { return kBandRows; }
