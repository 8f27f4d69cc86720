// Internal (non-ABI) declarations shared between the replay and learner translation units.
#pragma once
#include "dz_common.cuh"

namespace dz {

// Optional per-batch outputs the fused learner path wants straight from the sampler:
// row pointers into the replay store (the gather is fused into the conv1 operand load) and the
// float32/int32 scalars exactly as they enter jit(update).
struct BatchExtras {
  const uint8_t** d_s_tm1_rows;
  const uint8_t** d_s_t_rows;
  int32_t* d_a;
  float* d_r;
  float* d_disc;
  float* d_w;
  int fused;
};

int launch_sample(const dz_replay_view* view, int prioritized, const dz_sample_inputs* in, const dz_sample_outputs* out,
                  int batch, const BatchExtras& ex, void* stream);
int launch_update_priorities(const dz_replay_view* view, const int64_t* d_indices, const float* d_priorities, int n,
                             double alpha, int64_t size, void* stream);

struct TcBatch;
int launch_tc(const char* tag, const TcBatch& tb, int bnj, void* stream);

}  // namespace dz
