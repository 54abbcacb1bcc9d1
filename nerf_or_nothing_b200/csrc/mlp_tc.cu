// mlp_tc.cu — tcgen05/TMEM MLP engine (placeholder until gemm_tc.cu lands in this round).
#include "mlp.cuh"
namespace nerf {
MlpEngine* make_tc_mlp(bool) { return nullptr; }
}  // namespace nerf
