// attn_qktv.cu — K3/K4 placeholder translation unit (tcgen05 kernel lands in a later commit).
#include "sdf_common.cuh"
using namespace sdf;
extern "C" int sdf_attn_qktv_fwd(const sdf_attn_qktv_fwd_args* a) {
  (void)a;
  set_error("sdf_attn_qktv_fwd: not built yet");
  return SDF_ERR_UNSUPPORTED;
}
extern "C" int sdf_attn_qktv_bwd(const sdf_attn_qktv_bwd_args* a) {
  (void)a;
  set_error("sdf_attn_qktv_bwd: not built yet");
  return SDF_ERR_UNSUPPORTED;
}
