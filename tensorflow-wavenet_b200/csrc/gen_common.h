// Arguments shared by the two fast-generation kernels (generator.cu: throughput mode, one CTA per
// stream group; generator_lat.cu: latency mode, one stream spread over the whole GPU).
#pragma once
#include <stdint.h>

#include "../../include/wavenet_b200.h"

namespace wn {

struct GenArgs {
  int L, C, S, Q, G, use_biases;
  int gc_card;                     // rows of the embedding table; ids outside [0, gc_card) contribute nothing
  int sum_d;                       // sum of dilations (ring rows per stream)
  int streams, n_steps, commit;
  float temperature;
  const float *causal, *filter, *gate, *dense, *skip, *gc_filter, *gc_gate, *filter_bias, *gate_bias,
      *dense_bias, *skip_bias, *post1, *post2, *post1_bias, *post2_bias, *gc_embedding;
  int32_t* hdr;                    // [streams][4]: prev_id, step, pending_id, pending_valid
  float* pending;                  // [streams][L][C] layer inputs of an uncommitted single step
  float* rings;                    // [streams][sum_d][C]
  const int32_t *inputs, *forced, *gc_ids;
  const double* uniforms;
  int32_t* samples_out;
  float* proba_out;
  int dil[WN_MAX_LAYERS];
  int ring_off[WN_MAX_LAYERS];     // row offset of each layer's ring
};

// latency-mode kernel (generator_lat.cu).  comm: device scratch of gen_lat_comm_bytes(), launch_seq: a counter
// that differs between consecutive launches on the same state (tags stale words of earlier launches as invalid).
int64_t gen_lat_comm_bytes(const wn_config* cfg);
bool gen_lat_eligible(const GenArgs& a);
void set_gen_timeline(long long* p);   // debug: 16 int64 %globaltimer stamps of one step
int gen_lat_run(const GenArgs& a, void* comm, uint32_t launch_seq, cudaStream_t st);
// pipelined form of the same kernel: 2 .. gen_pipe_max_streams() streams share the layer-per-warp chain
bool gen_pipe_eligible(const GenArgs& a);
int gen_pipe_max_streams();
int gen_pipe_run(const GenArgs& a, void* comm, uint32_t launch_seq, cudaStream_t st);

}  // namespace wn
