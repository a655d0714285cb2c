// BF16 tensor-core path (placeholder until the tcgen05 kernels land).
#include "model.h"

namespace vapb {
int bf16_prepare(Model&) { return 0; }
void bf16_release(Model&) {}
size_t workspace_bytes_bf16(const Model&, const Geometry&) { return 0; }
int forward_bf16(Model& m, cudaStream_t, const float*, const Geometry&, char*, float*, float*, float*, const float**) {
  m.err = "bf16 mode not built";
  return -5;
}
int stage_bf16(const Model&, const Geometry&, char*, const std::string&, StageRef*) { return -1; }
}  // namespace vapb
