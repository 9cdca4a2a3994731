// sm_100a tcgen05/TMA fast path of the IPA layer (placeholder until the kernel lands).
#include "common.cuh"

extern "C" {
size_t dab_ipa_packed_bytes(const DabIpaDims*) { return 0; }
int dab_ipa_pack_weights(const DabIpaDims*, const DabIpaWeights*, void*, void*) {
  dab::set_error("dab_ipa_pack_weights: sm_100a fast path not built");
  return DAB_EUNSUPPORTED;
}
size_t dab_ipa_sm100_workspace_bytes(const DabIpaDims*) { return 0; }
int dab_ipa_fwd_sm100(const DabIpaDims*, const void*, const float*, const void*, const float*, const float*, float*,
                      void*, size_t, void*) {
  dab::set_error("dab_ipa_fwd_sm100: sm_100a fast path not built");
  return DAB_EUNSUPPORTED;
}
}
