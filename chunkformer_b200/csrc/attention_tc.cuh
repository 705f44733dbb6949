// tcgen05 chunk-pair attention kernel (placeholder until the kernel lands).
#pragma once
#include <string>

#include "attention_simt.cuh"

namespace cf {
constexpr bool kAttentionTcReady = false;
inline bool launch_attention_tc(const AttnParams&, cudaStream_t, std::string* err) {
  if (err) *err = "attention: tcgen05 kernel not built";
  return false;
}
}  // namespace cf
